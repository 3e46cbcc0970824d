#!/usr/bin/env python
"""bench.py -- one PPBO iteration (GP Laplace fit + RFF acquisition over the xi-grids of every query direction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config ackley20d] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full iteration on synthetic inputs of the named BASELINE.json shape (default: Ackley-20D, 200 queries x
m = 25 -> N = 5200 rows / 5000 pseudo-observations, F = 1000 features, 20 directions x 1024 grid points, 32768 samples):
  value : ms per iteration, inputs already resident in HBM (max over ranks, CUDA events)
  e2e   : the same iteration through the public host API with pinned HOST buffers: H2D of the inputs and D2H of the
          per-direction sums inside the timed region, followed by the host arg-max over directions
Only the Monte-Carlo samples are partitioned over the ranks (fixed total work -> "scaling": "strong").
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ppbo_iteration_ms"
UNIT = "ms"
# FP64 tensor (DMMA) pipe peak measured on this pool's B200 by scratch/ubench.cu (profiles/r01_fp64_peaks_ubench.txt):
# register-resident mma.sync.m8n8k4.f64 loop, 37.1 TFLOP/s; MEASURED_PEAKS.json carries no FP64 figure.
FP64_TENSOR_PEAK_TFLOPS = 37.1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="ackley20d")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--samples", type=int, default=None, help="override S (debug)")
    ap.add_argument("--profile", action="store_true", help="1 warm-up + --steps resident steps, no JSON (for ncu)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_sample(prob, fit_iters_full=None, q_sample=60, s_sample=512, threads=None):
    """The reference algorithm (oracle port, numpy/scipy + multi-threaded BLAS) on a BOUNDED sample of the workload,
    extrapolated to one full iteration in ms.  What is timed and how it is scaled is returned in `sample`.

    Reference recipe per iteration: Sigma = reg(K) ; Sigma^-1 (dposv) ; trust-exact on the dense N x N
    Hessian (one T_hessian + >= 1 dpotrf per outer iteration, ~100-150 outer iterations from the reference's random
    start, probe in SURVEY.md 3.2) ; posterior covariance (dposv) ; RFF: Phi_X, weight-space trust-exact, Omega,
    Phi_grid, Fs = Omega Phi_grid, per-sample max.  N^3 pieces are timed at N_s = q_sample (m+1) rows and scaled by
    (N/N_s)^3, N^2 pieces by (N/N_s)^2, the sampling GEMM at s_sample samples and scaled by S/s_sample."""
    from oracle import ppbo_oracle as O
    import scipy.linalg
    D, Q, m, S, P, F = prob["D"], prob["Q"], prob["m"], prob["S"], prob["P"], prob["F"]
    theta = prob["theta"]
    qs = min(Q, q_sample)
    Ns, N = qs * (m + 1), Q * (m + 1)
    Xs = prob["X"][:Ns]
    c3, c2 = (N / Ns) ** 3, (N / Ns) ** 2
    parts = {}

    def timed(name, scale, fn):
        t0 = time.perf_counter()
        out = fn()
        parts[name] = (time.perf_counter() - t0) * scale
        return out
    Sigma = timed("gram_regularize_closed_form", c2, lambda: O.regularize_covariance(O.se_kernel(Xs, Xs, theta), svd_roundtrip=False))
    Sinv = timed("Sigma_inverse_dposv", c3, lambda: O.pd_inverse(Sigma))
    f = np.zeros(Ns)
    outer = 120 if fit_iters_full is None else fit_iters_full      # reference outer iterations (95-156)

    def one_outer():
        H = -O.T_hessian(f, Sinv, qs, m, theta[0])
        g = -O.T_grad(f, Sinv, qs, m, theta[0])
        c, low = scipy.linalg.cho_factor(H, lower=True)
        return scipy.linalg.cho_solve((c, low), -g) + O.T_value(f, Sinv, qs, m, theta[0], quadrature=False)
    timed("trust_exact_outer_iterations(x%d)" % outer, c3 * outer, one_outer)
    timed("posterior_covariance_dposv", c3, lambda: O.pd_inverse(Sinv - O.create_Lambda(f, qs, m, theta[0])))
    if F:
        W, b = prob["W"], prob["b"]
        PhiX = timed("rff_features_design", N / Ns, lambda: O.rff_features(W, b, Xs, theta[2]))
        w = np.zeros(F)
        timed("rff_trust_exact_outer_iterations(x30)", 30 * N / Ns, lambda: (O.rff_S_grad(w, PhiX, qs, m, theta[0]),
                                                                          O.rff_S_hess_diag(w, PhiX, qs, m, theta[0]),
                                                                          O.rff_S(w, PhiX, qs, m, theta[0], quadrature=False)))
        ss = min(S, s_sample)
        rng = np.random.RandomState(0)
        Omega = rng.randn(ss, F)

        def sampling():
            out = 0.0
            for d in range(prob["grids"].shape[0]):
                Phi = O.rff_features(W, b, prob["grids"][d], theta[2])
                mx, _ = O.rff_eval_argmax(Omega, Phi)
                out += np.maximum(mx, 0).sum()
            return out
        timed("rff_sampling_gemm_rowmax", S / ss, sampling)
    total_ms = 1e3 * sum(parts.values())
    sample = ("oracle port (numpy/scipy, closed-form regularisation, GH quadrature replaced by ndtr) timed at N_s=%d rows "
              "(N^3 parts x%.1f, N^2 parts x%.1f), 1 trust-exact outer iteration x%d, %d of %d samples on all %d grids; "
              "extrapolated parts ms: %s" % (Ns, c3, c2, outer, min(S, s_sample), S, prob["grids"].shape[0],
                                             {k: round(1e3 * v, 1) for k, v in parts.items()}))
    return total_ms, sample


def int8_peak_tops():
    """INT8 dense tensor peak in TOP/s: 2 x MEASURED_PEAKS.json's cuBLAS bf16 burst figure (kernel timed alone), else 2 x the
    profiling guide's fallback of 1590 TFLOP/s."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            mp = json.load(fh)
        return 2.0 * float(mp["bf16_tflops"]), ("2 x MEASURED_PEAKS.json bf16_tflops (%.1f, burst; sustained %.1f): INT8 runs on the "
                                                "same tensor pipe at twice the K per instruction" % (mp["bf16_tflops"],
                                                                                                      mp.get("bf16_tflops_sustained", float("nan"))))
    except Exception:
        return 2.0 * 1590.0, "2 x 1590 TFLOP/s bf16 (B200_PROFILING.md fallback; MEASURED_PEAKS.json absent)"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference algorithm on the host cores (oracle port; the Python reference cannot travel)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ppbo_b200 import synthetic
    prob = synthetic.make_problem(args.config, S=args.samples)
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample = cpu_reference_sample(prob)
        if i >= args.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(prob, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def iteration_slices():
    from ppbo_b200 import iteration
    return iteration.SAMPLING_SLICES


def workload_config(prob, n_gpus):
    return {"workload": "%s: D=%d, Q=%d queries x m=%d (N=%d rows, %d pseudo-observations), %s theta=%s, F=%d RFF features, "
                        "%d query directions x P=%d xi-grid points, S=%d samples" % (
                            prob["name"], prob["D"], prob["Q"], prob["m"], prob["N"], prob["Q"] * prob["m"], prob["kernel"],
                            prob["theta"], prob["F"], prob["grids"].shape[0], prob["P"], prob["S"]),
            "parallelism": ("GP fit + weight-space fit (background thread) + all S samples on one GPU" if n_gpus == 1 else
                            "GP fit on rank 0 (no samples: it is the critical path), weight-space fit on rank 1, broadcast of (omega_MAP, "
                            "diag Hessian); S sharded over ranks 1..%d, which sample while rank 0 fits; broadcast of mu*; one "
                            "all-reduce of 3 x directions doubles" % (n_gpus - 1)),
            "l2_policy": "working set per step (Sigma, G, factor, Omega, PhiT: > 1 GB) exceeds the 126 MB L2; no explicit flush",
            "fit_start": "cold: GP Newton from f = 0; weight-space Newton from omega = 0, concurrently with the GP fit",
            "sampling_engine": "tcgen05 INT8, %d digit planes per operand (error-free splitting; FP64-GEMM accuracy)" % iteration_slices()}


# ------------------------------------------------------------------------------------------------- our arm
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from ppbo_b200 import _lib, iteration, ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    shard = iteration.Shard()

    prob = synthetic.make_problem(args.config, S=args.samples)
    kernel, theta, Q, m, S = prob["kernel"], prob["theta"], prob["Q"], prob["m"], prob["S"]
    B, P, D = prob["grids"].shape
    Fdim = prob["F"]
    # f_init = None: cold start of the GP Newton iteration at f = 0; omega0 = None: weight-space start projected from the GP mode
    inputs = iteration.IterationInputs(prob["X"], None, prob["W"], prob["b"], None, prob["grids"])
    sums_host = torch.empty((B, 3), dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gemm_events = []

    def step(resident=None, timers=None):
        d = resident if resident is not None else inputs.to_device(dev)
        sums, gp, rff = iteration.run_iteration(d, kernel, theta, Q, m, S, shard=shard, seed=1234, timers=timers)
        if resident is None:
            sums_host.copy_(sums, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            ei, var = iteration.acquisition_values(sums_host.numpy(), S)
            return int(np.argmax(ei)), gp, rff
        return sums, gp, rff

    step_marks = {}
    held = [None]

    def timed_run(resident):
        import gc
        gc.collect()
        gc.disable()                       # a generation-2 collection inside the timed region is a host hiccup, not the workload
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.ppbo_launch_count()
        marks = [time.perf_counter()]
        e0.record()
        out, held[0] = held[0], None       # the previous step's products stay alive until the next step has replaced them,
        for _ in range(args.steps):        # in the warm-up as in the timed steps: the allocator pool is in its steady state
            out = step(resident)
            marks.append(time.perf_counter())      # host time after the step was issued (e2e: after its result arrived)
        held[0] = out
        e1.record()
        torch.cuda.synchronize()
        gc.enable()
        ms = e0.elapsed_time(e1)
        launches = lib.ppbo_launch_count() - l0
        step_marks["resident" if resident is not None else "e2e"] = [1e3 * (b - a) for a, b in zip(marks[:-1], marks[1:])]
        barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps, launches, out

    resident = inputs.to_device(dev)
    if args.profile:
        for _ in range(1 + args.steps):
            step(resident)
        torch.cuda.synchronize()
        return
    for _ in range(max(args.warmup, 3)):
        held[0] = step(resident)           # (a warm-up that drops its products at once leaves the second timed step to
    clocks = ClockSampler(local)           # cudaMalloc a second set of 200 MB buffers: +1 ms on one GPU, up to +10 ms on two)
    if rank == 0:
        clocks.start()
    ms_dev, launches, out = timed_run(resident)
    fit_iters = out[1].lap.stats["iterations"] if out[1] is not None else None
    rff_iters = out[2].stats["iterations"] if (out[2] is not None and out[2].stats) else None
    for _ in range(2):
        held[0] = step(None)
    ms_e2e, _, out2 = timed_run(None)
    clk = clocks.stop() if rank == 0 else None

    # stage breakdown + the dominant kernel (sampling GEMM with fused row max) timed with CUDA events on its stream
    barrier()
    stage_ms, gemm_ms = {}, []
    for _ in range(max(3, min(args.steps, 5))):
        timers = []
        step(resident, timers)
        torch.cuda.synchronize()
        for (n0, a), (n1, b_) in zip(timers[:-1], timers[1:]):
            stage_ms.setdefault(n1, []).append(a.elapsed_time(b_))
    lo, hi = shard.sample_bounds(S)
    if hi == lo:                      # rank 0 of several takes no samples: time the launch a sampling rank performs
        lo, hi = 0, S // max(1, world - 1)
    rffs = out[2]
    Omega = ops.rff_sample_omega(rffs.omega_map, rffs.hess_diag, hi - lo, seed=1234, sample0=lo)
    PhiT = iteration.rff_grid_features(resident["W"], resident["b"], theta[2], resident["grids"])
    engine = iteration.sampling_engine(hi - lo, P, Fdim)
    ks = iteration.SAMPLING_SLICES
    flops = 2.0 * (hi - lo) * Fdim * P * B                       # 2 S F P per direction
    if engine == "i8":
        ap, asc = ops.ozaki_slice(Omega, 0, ks)
        bp, bsc = ops.ozaki_slice(PhiT, 1, ks)
        fm = torch.empty((B, hi - lo), dtype=torch.float64, device=dev)
        am = torch.empty((B, hi - lo), dtype=torch.int32, device=dev)

        def dominant():
            ops.ozaki_rowmax(ap, asc, hi - lo, bp, bsc, P, B, Fdim, ks, fmax=fm, arg=am)
    else:
        def dominant():
            ops.rff_eval_argmax(Omega, PhiT)
    for i in range(3 + 5):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dominant()
        b_.record()
        torch.cuda.synchronize()
        if i >= 3:
            gemm_ms.append(a.elapsed_time(b_))
    gemm_t = float(np.mean(gemm_ms))
    fp64_equiv = flops / (gemm_t * 1e-3) / 1e12
    if engine == "i8":
        # KS (KS+1)/2 exact INT8 plane products per FP64 product; INT8 dense peak = 2 x the measured bf16 peak (same tensor
        # pipe, twice the K per instruction: nominal 4.5 POP/s against 2.25 PFLOP/s)
        ops_per_launch = flops * ks * (ks + 1) / 2
        achieved = ops_per_launch / (gemm_t * 1e-3) / 1e12
        peak, peak_src = int8_peak_tops()
        roof = {"kernel": "ozaki_rowmax_kernel<%d,64> (RFF sampling contraction on tcgen05.mma.kind::i8, %d digit planes per operand, "
                          "fused per-sample max/arg-max)" % (ks, ks),
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape (profiles/r01_ncu_ozaki_rowmax.txt);
                # algorithmic HBM bytes: the A digit planes once per grid + the B planes = 20 x 201 MB + 126 MB = 4.15 GB
                "traffic": (4.501766e9 + 2.0655e7) if (args.config == "ackley20d" and hi - lo == 32768 and ks == 6 and P == 1024) else None,
                "traffic_unit": "bytes per launch (ncu)",
                "kernel_ms": gemm_t, "ops_per_launch": ops_per_launch, "fp64_equivalent_tflops": fp64_equiv,
                "fp64_dmma_peak_tflops": FP64_TENSOR_PEAK_TFLOPS, "peak_source": peak_src}
    else:
        achieved = fp64_equiv
        roof = {"kernel": "gemm_nt_rowmax_kernel (RFF sampling GEMM, fused per-sample max/arg-max)", "bound": "tensor",
                "achieved": achieved, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_TENSOR_PEAK_TFLOPS,
                "traffic": None, "kernel_ms": gemm_t, "flops_per_launch": flops,
                "peak_source": "FP64 DMMA pipe measured on this pool (profiles/r01_fp64_peaks_ubench.txt); "
                               "MEASURED_PEAKS.json has no FP64 figure (cuBLAS DGEMM 8192^3 reaches 35.5)"}
    sample_points_per_s = S * P * B / (float(np.mean(stage_ms["acquisition"])) * 1e-3) if shard.world == 1 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": ms_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(prob, world),
        "e2e": {"value": ms_e2e, "unit": UNIT, "h2d_bytes_per_step": inputs.nbytes(), "d2h_bytes_per_step": B * 3 * 8,
                "selected_direction": out2[0]},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "stages_ms": {k: float(np.mean(v)) for k, v in stage_ms.items()},
        # diagnostics: host time per issued step (resident steps are issued asynchronously, e2e steps end with the result on the host)
        "resident_host_ms_per_step": [round(t, 2) for t in step_marks.get("resident", [])],
        "e2e_host_ms_per_step": [round(t, 2) for t in step_marks.get("e2e", [])],
        "fit_newton_iterations": fit_iters, "rff_newton_iterations": rff_iters,
        "rff_sample_points_per_s": sample_points_per_s,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, sample = cpu_reference_sample(prob)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
