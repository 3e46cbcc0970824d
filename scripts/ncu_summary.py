"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/ (run here, no GPU needed).
  python scripts/ncu_summary.py launches gpurun_out/r01_launches.csv            -> per-kernel share table
  python scripts/ncu_summary.py rep gpurun_out/r01_rowmax.ncu-rep               -> key raw metrics per captured launch
  python scripts/ncu_summary.py sass ppbo_b200/libppbo_b200.so                  -> per-kernel registers / spills / shared memory
                                                                                   and counts of the Blackwell SASS mnemonics"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        # tcgen05 kernels (csrc/ozaki.cu)
        "gpc__cycles_elapsed.max", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_writes.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_uniform_realtime.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r["Metric Unit"], v)
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("# %d launches, %.1f us total (cold-cache, serialised: compare SHARES, not absolutes)" % (sum(len(v) for v in agg.values()), tot))
    print("%-78s %6s %12s %9s %7s" % ("kernel", "n", "total_us", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-78s %6d %12.1f %9.1f %6.1f%%" % (k[:78], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))


def rep(path, keys=KEYS):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = [re.sub(r"^[A-Z_]+\.Triage[A-Za-z]*\.", "", h) for h in rows[0]], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== %s  grid=%s block=%s" % (re.sub(r"\(.*", "", r[col["Kernel Name"]])[:100], r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
        for k in keys:
            if k in col:
                print("   %-70s %s %s" % (k, r[col[k]], units[col[k]]))


def traffic(path, key, out_json="profiles/traffic.json", pattern=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of the (first matching) captured launch -> profiles/traffic.json[key]"""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if pattern and pattern not in r[col["Kernel Name"]]:
            continue
        def val(name):
            v = float(r[col[name]].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[col[name]], 1)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        rec = json.load(open(out_json)) if os.path.exists(out_json) else {}
        rec[key] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "kernel": r[col["Kernel Name"]][:120],
                    "duration_ns_under_ncu": r[col["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in col else None,
                    "source": "ncu --set full of " + os.path.basename(path)}
        json.dump(rec, open(out_json, "w"), indent=1)
        print(key, rec[key])
        return
    print("no matching launch in", path)


MNEMONICS = ["UTCIMMA", "UTCCP", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "DMMA", "LDGSTS", "REDUX", "ATOM", "RED."]


def sass(path):
    """static evidence from the built library (cuobjdump, no GPU): which kernels issue tcgen05 (UTCIMMA / UTCCP / LDTM / UTCBAR),
    TMA bulk copies (UBLKCP) with mbarrier waits (SYNCS), FP64 tensor-core MMAs (DMMA), cp.async (LDGSTS); registers and spills"""
    def demangle(names):
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, [re.sub(r"^void ", "", re.sub(r"\(.*", "", d)) for d in out]))
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", path], capture_output=True, text=True).stdout
    usage = {m.group(1): m.group(2) for m in re.finditer(r"Function (\S+):\n\s+(REG:.*)", res)}
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        cur["_inst"] += 1
        for k in MNEMONICS:
            if k in line:
                cur[k] += 1
    spills = {}
    log = os.path.join(os.path.dirname(os.path.abspath(path)), "build", "ptxas.log")     # written by ppbo_b200/build.py (-Xptxas -v)
    if os.path.exists(log):
        for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads",
                             open(log).read()):
            spills[m.group(1)] = (int(m.group(3)), int(m.group(4)))
    names = demangle(list(counts))
    print("# %s: %d kernels, arch sm_100a; columns after smem are SASS instruction counts (static, not executed)" % (path, len(counts)))
    print("# spill = bytes of spill stores/loads per thread reported by ptxas -v (stack frames without spills are local arrays and the "
          "slow paths of sin / cos / exp / erfc)")
    print("%-64s %4s %5s %7s %7s %6s  %s" % ("kernel", "reg", "stack", "spill", "smem", "inst", "mnemonics"))
    tot = collections.Counter()
    for k, c in counts.items():
        u = dict(f.split(":") for f in usage.get(k, "").split() if ":" in f)
        tot.update(c)
        marks = " ".join("%s=%d" % (m, c[m]) for m in MNEMONICS if c[m])
        print("%-64s %4s %5s %7s %7s %6d  %s" % (names[k][:64], u.get("REG", "?"), u.get("STACK", "?"), "%d/%d" % spills.get(k, (0, 0)),
                                                 u.get("SHARED", "?"), c["_inst"], marks))
    print("# totals: " + " ".join("%s=%d" % (m, tot[m]) for m in MNEMONICS))
    print("# kernels with register spills: " + (", ".join("%s (%d/%d B)" % ((names[k][:60],) + spills[k]) for k in counts
                                                          if sum(spills.get(k, (0, 0)))) or "none"))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:])
    else:
        {"launches": launches, "rep": rep, "sass": sass}[sys.argv[1]](sys.argv[2])
