"""GPU probe: cold GP fit (with / without the factor at the mode) against the incremental path (GPState.append), the warm
weight-space fit and the single-point posterior mean.  Prints stage times (CUDA events) and the fit statistics."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from ppbo_b200 import iteration, ops, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ackley20d"
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = synthetic.CONFIGS[name]
prob = synthetic.make_problem(name, Q=cfg["Q"] + extra)
m, Q0, theta = prob["m"], cfg["Q"], prob["theta"]
dev = torch.device("cuda", 0)
X = ops.to_dev(prob["X"])
W, b = ops.to_dev(prob["W"]), ops.to_dev(prob["b"])
n0 = Q0 * (m + 1)


def timed(fn, reps=3):
    out, ts = None, []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    return out, ts


for fam in (False, True):
    g, ts = timed(lambda: iteration.gp_fit(X[:n0], prob["kernel"], theta, Q0, m, tol=1e-8, factor_at_mode=fam))
    print("cold gp_fit factor_at_mode=%s: %s ms  stats %s" % (fam, ["%.2f" % t for t in ts], g.lap.stats))

st = iteration.GPState(prob["kernel"], theta, prob["D"], m, Q0 + extra, dev, tol=1e-8)
_, ts = timed(lambda: st.cold(X[:n0]), reps=2)
print("GPState.cold: %s ms" % ["%.2f" % t for t in ts])
rs = iteration.RFFState(W, b, theta, m, Q0 + extra, tol=1e-8)
_, ts = timed(lambda: rs.cold(X[:n0]), reps=2)
print("RFFState.cold: %s ms  stats %s" % (["%.2f" % t for t in ts], rs.fit.stats))
for q in range(Q0, Q0 + extra):
    blk = X[q * (m + 1):(q + 1) * (m + 1)]
    torch.cuda.synchronize()
    a, e, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    st.append(blk)
    e.record()
    rs.append(blk)
    e2.record()
    torch.cuda.synchronize()
    s = st.lap.stats
    print("append -> Q=%d: gp %.2f ms (its %d, chord %d, factorizations %d, warm_first_rel %.2e, last_rel %.1e, conv %d) | rff %.2f ms %s" % (
        q + 1, a.elapsed_time(e), s["iterations"], s["chord_steps"], s["factorizations"], s["warm_first_rel"], s["last_rel_step"],
        s["converged"], e.elapsed_time(e2), rs.fit.stats))
# check against a cold fit at the final size
g = iteration.gp_fit(X, prob["kernel"], theta, Q0 + extra, m, tol=1e-9)
f1, f2 = st.f_map.cpu().numpy(), g.f_map.cpu().numpy()
print("warm vs cold mode: rel %.2e" % (np.abs(f1 - f2).max() / np.abs(f2).max()))
pm = ops.PointMean(prob["kernel"], X, theta[1], theta[2], g.alpha)
x = np.random.RandomState(0).rand(prob["D"])
pm(x)
t0 = time.perf_counter()
for _ in range(2000):
    pm(x)
print("mu_pred_point: %.1f us per evaluation (N = %d)" % ((time.perf_counter() - t0) / 2000 * 1e6, X.shape[0]))
