"""Gram / posterior-mean timings at the Ackley-20D sizes: tensor-pipe kernels against the difference-form kernels (tuning key 11)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppbo_b200 import _lib, ops, synthetic, iteration
lib = _lib.load()
prob = synthetic.make_problem("ackley20d")
X = ops.to_dev(prob["X"]); th = prob["theta"]; Q, m = prob["Q"], prob["m"]
grids = ops.to_dev(prob["grids"]); B, P, D = grids.shape
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
out = {}
for key in (1, 0):
    lib.ppbo_set_tuning(11, key)
    name = "difference form" if key else "tensor pipe"
    t = timeit(lambda: ops.gram_regularized("SE_kernel", X, th[1], th[2], 1e-6))
    S = ops.gram_regularized("SE_kernel", X, th[1], th[2], 1e-6)
    N = X.shape[0]
    print("%-16s gram N=%d: %.3f ms (%.0f GB/s write)" % (name, N, t, 8 * N * N / t / 1e6))
    Kx = ops.kernel_matrix("SE_kernel", grids.reshape(B * P, D)[:4096], X, th[1], th[2])
    t = timeit(lambda: ops.kernel_matrix("SE_kernel", grids.reshape(B * P, D)[:4096], X, th[1], th[2]))
    print("%-16s cross 4096 x %d: %.3f ms (%.0f GB/s write)" % (name, N, t, 8 * 4096 * N / t / 1e6))
    out[key] = (S.clone(), Kx.clone())
lib.ppbo_set_tuning(11, 0)
print("gram   max |new - old| / max: %.2e   symmetric: %s" % (float((out[0][0] - out[1][0]).abs().max() / out[1][0].abs().max()), bool(torch.equal(out[0][0], out[0][0].T))))
print("cross  max |new - old| / max: %.2e" % float((out[0][1] - out[1][1]).abs().max() / out[1][1].abs().max()))
Sigma = out[0][0]
fit = ops.laplace_fit(Sigma, Q, m, th[0], tol=1e-8)
g = iteration.GPFit(); g.X, g.kernel, g.theta, g.Q, g.m, g.lengthscales, g.Sigma, g.lap = X, "SE_kernel", th, Q, m, th[1], Sigma, fit
cand = grids.reshape(B * P, D)
t = timeit(lambda: iteration.mustar_over_candidates(g, cand))
mu = iteration.posterior_mean(g, cand)
print("tensor pipe      mu* over %d candidates: %.3f ms" % (cand.shape[0], t))
# against the materialised cross-covariance of the difference-form kernel (library exp)
lib.ppbo_set_tuning(11, 1)
Kc = ops.kernel_matrix("SE_kernel", cand[:4096], X, th[1], th[2])
lib.ppbo_set_tuning(11, 0)
ref = Kc @ fit.alpha
print("mean   max |mat-vec mode - K alpha| / max: %.2e" % float((mu[:4096] - ref).abs().max() / ref.abs().max()))
