"""Concurrency timeline of one PPBO iteration (diagnostics; there is no nsys in the image).

torch.profiler's CUPTI activity records see every kernel of the process, including those the C-ABI library launches, with
stream, start and duration.  This script profiles cold and appended iterations of the bench problem and prints, per
iteration: the busy time of every stream, the union busy time, and a coarse text timeline (0.25 ms buckets: which streams
had a kernel running, and the kernel that covered most of the bucket).

    python scripts/timeline.py [--mode steady|cold] [--steps 2] [--out gpurun_out/timeline_steady.txt]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppbo_b200 import iteration, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="steady")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("--bucket", type=float, default=0.25)
    ap.add_argument("--api", action="store_true", help="also record the runtime API calls: the dump then shows how long after its launch call "
                    "each kernel started (host-side or device-side delay?)")
    ap.add_argument("--dump", default=None, help="file for the full kernel sequence (stream, start us, duration us, name) of the last iteration")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    name = "ackley20d"
    cfg = synthetic.CONFIGS[name]
    Q0, extra = cfg["Q"], 4 + args.steps
    big = synthetic.make_problem(name, Q=Q0 + extra)
    prob = synthetic.make_problem(name)
    kernel, theta, m, S = prob["kernel"], prob["theta"], prob["m"], prob["S"]
    inputs = iteration.IterationInputs(prob["X"], None, prob["W"], prob["b"], None, prob["grids"])
    res = inputs.to_device(dev)
    B, P, D = prob["grids"].shape
    blocks = [torch.from_numpy(np.ascontiguousarray(big["X"][(Q0 + i) * (m + 1):(Q0 + i + 1) * (m + 1)])).to(dev) for i in range(extra)]
    state = iteration.IterationState(kernel, theta, D, m, Q0 + extra, dev, res["W"], res["b"])

    def cold():
        return iteration.run_iteration(res, kernel, theta, Q0, m, S, seed=1)

    def steady(i):
        d = {"block": blocks[i], "W": res["W"], "b": res["b"], "grids": res["grids"]}
        return iteration.run_iteration(d, kernel, theta, Q0 + i + 1, m, S, seed=1, state=state)

    if args.mode == "gpfit":            # the cold GP fit alone (one stream of the caller + the library's Cholesky streams)
        from ppbo_b200 import ops
        for _ in range(3):
            g = iteration.gp_fit(res["X"], kernel, theta, Q0, m, tol=1e-8)
        torch.cuda.synchronize()
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            g = iteration.gp_fit(res["X"], kernel, theta, Q0, m, tol=1e-8)
            torch.cuda.synchronize()
        prof.export_chrome_trace("/tmp/ppbo_trace.json")
        ev = sorted((e for e in json.load(open("/tmp/ppbo_trace.json"))["traceEvents"] if e.get("cat") == "kernel"), key=lambda e: e["ts"])
        t0 = ev[0]["ts"]
        with open(args.dump or "gpurun_out/seq_gpfit.txt", "w") as fh:
            fh.write("# cold GP fit alone: %.2f ms, %d kernels, stats %s\n" % ((max(e["ts"] + e["dur"] for e in ev) - t0) / 1e3, len(ev), g.lap.stats))
            for e in ev:
                fh.write("s%-4d %9.1f %8.1f  %s grid=%s\n" % (e["args"]["stream"], e["ts"] - t0, e["dur"], short(e["name"]), e["args"].get("grid", "")))
        return
    keep = []
    if args.mode == "cold":
        for _ in range(3):
            keep.append(cold())
    else:
        iteration.run_iteration(res, kernel, theta, Q0, m, S, seed=1, state=state)
        for i in range(3):
            keep.append(steady(i))
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    marks = []
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU] if args.api else [ProfilerActivity.CUDA]) as prof:
        for k in range(args.steps):
            torch.cuda.synchronize()
            keep.append(cold() if args.mode == "cold" else steady(3 + k))
            torch.cuda.synchronize()
    tmp = "/tmp/ppbo_trace.json"
    prof.export_chrome_trace(tmp)
    allev = json.load(open(tmp))["traceEvents"]
    ev = [e for e in allev if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    launch_ts = {e["args"]["correlation"]: (e["ts"], e.get("tid"), e.get("dur", 0)) for e in allev
                 if e.get("cat") == "cuda_runtime" and "correlation" in e.get("args", {})}
    out = open(args.out, "w") if args.out else sys.stdout
    # an iteration ends with its acq_reduce_kernel (the reduction of the sampled maxima)
    steps, cur = [], []
    for e in ev:
        cur.append(e)
        if "acq_reduce_kernel" in e["name"]:
            steps.append(cur)
            cur = []
    if cur and steps:
        steps[-1].extend(cur)
    elif cur:
        steps.append(cur)
    for si, st in enumerate(steps):
        t0 = st[0]["ts"]
        t1 = max(e["ts"] + e["dur"] for e in st)
        streams = sorted({e["args"]["stream"] for e in st})
        out.write("== %s iteration %d: %.2f ms, %d kernels, streams %s\n" % (args.mode, si, (t1 - t0) / 1e3, len(st), streams))
        for s_ in streams:
            k = [e for e in st if e["args"]["stream"] == s_]
            busy = sum(e["dur"] for e in k)
            out.write("   stream %3d: %4d kernels, busy %.2f ms, first %.2f ms, last end %.2f ms, top: %s\n" % (
                s_, len(k), busy / 1e3, (k[0]["ts"] - t0) / 1e3, (max(e["ts"] + e["dur"] for e in k) - t0) / 1e3,
                ", ".join("%s x%d %.2fms" % (n[:28], c, d / 1e3) for n, c, d in top(k))))
        # union busy
        iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in st)
        u, ce = 0.0, iv[0][0]
        for a, b in iv:
            if b > ce:
                u += b - max(a, ce)
                ce = b
        out.write("   union busy %.2f ms of %.2f\n" % (u / 1e3, (t1 - t0) / 1e3))
        nb = int(np.ceil((t1 - t0) / 1e3 / args.bucket))
        for bi in range(nb):
            lo, hi = t0 + bi * args.bucket * 1e3, t0 + (bi + 1) * args.bucket * 1e3
            row = []
            for s_ in streams:
                cov = {}
                for e in st:
                    if e["args"]["stream"] != s_:
                        continue
                    o = min(hi, e["ts"] + e["dur"]) - max(lo, e["ts"])
                    if o > 0:
                        cov[e["name"]] = cov.get(e["name"], 0) + o
                if cov:
                    n, c = max(cov.items(), key=lambda kv: kv[1])
                    row.append("s%d:%3d%% %s" % (s_, int(100 * sum(cov.values()) / (hi - lo)), short(n)))
            out.write("   %6.2f  %s\n" % (bi * args.bucket, " | ".join(row)))
    if args.out:
        out.close()
    if args.dump:
        with open(args.dump, "w") as fh:
            for si, st in enumerate(steps):
                t0 = st[0]["ts"]
                fh.write("# iteration %d\n" % si)
                for e in st:
                    g = e["args"].get("grid", "")
                    lt = launch_ts.get(e["args"].get("correlation"))
                    extra = "  launched %9.1f (+%.1f, call %.1f us) tid %s" % (lt[0] - t0, e["ts"] - lt[0], lt[2], lt[1]) if lt else ""
                    fh.write("s%-4d %9.1f %8.1f  %s grid=%s%s\n" % (e["args"]["stream"], e["ts"] - t0, e["dur"], short(e["name"]), g, extra))


def short(n):
    n = n.replace("ppbo::", "").replace("void ", "").replace("oz::", "")
    return n[:26]


def top(k):
    agg = {}
    for e in k:
        a = agg.setdefault(short(e["name"]), [0, 0.0])
        a[0] += 1
        a[1] += e["dur"]
    return sorted(((n, c, d) for n, (c, d) in agg.items()), key=lambda x: -x[2])[:4]


if __name__ == "__main__":
    main()
