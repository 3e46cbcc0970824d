"""GPU probe: sampling-contraction kernel time against the work-order parameters (tuning key 4 = column groups NG, key 12 = grids per
block GB) on the bench shape.  The kernel is power-limited, so the configurations are visited round-robin (every round in a different
order) and the per-configuration median is reported."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ppbo_b200 import _lib, iteration, ops, synthetic  # noqa: E402

lib = _lib.load()
prob = synthetic.make_problem("ackley20d")
dev = torch.device("cuda", 0)
S, F = prob["S"], prob["F"]
B, P, D = prob["grids"].shape
rng = np.random.RandomState(0)
Omega = ops.to_dev(rng.randn(S, F))
W, b = ops.to_dev(prob["W"]), ops.to_dev(prob["b"])
PhiT = iteration.rff_grid_features(W, b, prob["theta"][2], ops.to_dev(prob["grids"]))
ap, asc = ops.ozaki_slice(Omega, 0, 6)
bp, bsc = ops.ozaki_slice(PhiT, 1, 6)
fm = torch.empty((B, S), dtype=torch.float64, device=dev)
am = torch.empty((B, S), dtype=torch.int32, device=dev)
cfgs = [(2, 1), (2, 2), (2, 4), (2, 7), (2, 20), (1, 5), (1, 10), (4, 1), (4, 4)]
times = {c: [] for c in cfgs}
for rnd in range(6):
    order = list(cfgs)
    np.random.RandomState(rnd).shuffle(order)
    for NG, GB in order:
        lib.ppbo_set_tuning(4, NG)
        lib.ppbo_set_tuning(12, GB)
        for i in range(3):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.ozaki_rowmax(ap, asc, S, bp, bsc, P, B, F, 6, fmax=fm, arg=am)
            e.record()
            torch.cuda.synchronize()
            if i >= 1 and rnd >= 1:
                times[(NG, GB)].append(a.elapsed_time(e))
for c in cfgs:
    print("NG=%d GB=%2d: median %.3f ms  min %.3f  max %.3f" % (c[0], c[1], np.median(times[c]), np.min(times[c]), np.max(times[c])), flush=True)
lib.ppbo_set_tuning(4, 0)
lib.ppbo_set_tuning(12, 0)
