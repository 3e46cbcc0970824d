"""GPU probe: cold GP fits with the tuning key 13 settings (1 = no mixing, 0 / 3 = early chord phase with Anderson mixing)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ppbo_b200 import _lib, iteration, ops, synthetic  # noqa: E402

lib = _lib.load()
for name in sys.argv[1:] or ["ackley20d", "levy10d", "hartmann6d", "camel2d"]:
    prob = synthetic.make_problem(name)
    m, Q0, theta = prob["m"], prob["Q"], prob["theta"]
    X = ops.to_dev(prob["X"])
    ref = None
    for key in (1, 0):
        lib.ppbo_set_tuning(13, key)
        ts = []
        for rep in range(4):
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g = iteration.gp_fit(X, prob["kernel"], theta, Q0, m, tol=1e-8)
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        f = g.f_map.cpu().numpy()
        if ref is None:
            ref = f
        s = g.lap.stats
        print("%-10s key13=%d cold: %.2f ms (min %.2f) its %d chord %d fact %d conv %d  |f - f_plain| rel %.2e" % (
            name, key, np.median(ts), np.min(ts), s["iterations"], s["chord_steps"], s["factorizations"], s["converged"],
            np.abs(f - ref).max() / np.abs(ref).max()), flush=True)
    lib.ppbo_set_tuning(13, 0)
