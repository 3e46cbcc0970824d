"""TEST INFRASTRUCTURE ONLY -- full-size oracle fixtures for the synthetic BASELINE problems the bench and the `-m gpu`
parity tests run (ppbo_b200/synthetic.py: levy10d N = 2080, ackley20d N = 5200), and the measured cost record of the
reference's own algorithm at those sizes.

    python oracle/make_full_fixtures.py fixture ackley20d     # -> tests/golden/full_ackley20d.npz
    python oracle/make_full_fixtures.py record  ackley20d     # -> profiles/r02_reference_full_fit_ackley20d.json

`fixture`: everything is computed by oracle/ppbo_oracle.py (numpy/scipy, FP64) on the host, independently of the CUDA path:
  * mode of T: the reference's scipy trust-exact on (-T, -T_grad, -T_hessian) (src/gp_model.py:382-384) started at f = 0 (the
    bench's cold start), then tightened by Newton on the fixed point f = Sigma beta(f) (ppbo_oracle.fmap_tight) FROM THAT POINT;
  * mu* over the candidate set of the bench pipeline (design rows + all grid points);
  * weight-space mode: trust-exact with the reference's diagonal Hessian (src/random_fourier_sampler.py:124-132) from omega = 0,
    tightened by Newton with the exact Hessian; the reference's diagonal Laplace covariance at that point;
  * sampled acquisition on the first SLICE samples of the Philox stream (seed 1234, the bench's seed) in FP64: per-sample
    max / arg-max / gap to the runner-up on every grid, the three sums per grid and the selected direction.
`record`: runs the oracle's trust-exact fit at full N to convergence from a random start f ~ N(0, Sigma) (the reference's
default, src/gp_model.py:374) and stores the outer-iteration count, wall time per iteration and the cost of the other N^3
pieces; bench.py's reference arm multiplies its own per-iteration timing (measured live at full N) by this iteration count.
"""
import json
import os
import sys
import time

import numpy as np
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ppbo_oracle as O  # noqa: E402
from ppbo_b200 import synthetic  # noqa: E402

SLICE = 2048
SEED = 1234


def trust_exact_counted(Sigma_inv, Q, m, sigma, f0, gtol=None, maxiter=None):
    """oracle trust-exact with per-iteration wall times"""
    import scipy.optimize
    times, t_last = [], [time.perf_counter()]

    def cb(xk, *a):
        now = time.perf_counter()
        times.append(now - t_last[0])
        t_last[0] = now
    opts = {}
    if gtol is not None:
        opts["gtol"] = gtol
    if maxiter is not None:
        opts["maxiter"] = maxiter
    res = scipy.optimize.minimize(lambda f: -O.T_value(f, Sigma_inv, Q, m, sigma, quadrature=False), np.asarray(f0, float).ravel(),
                                  method="trust-exact", jac=lambda f: -O.T_grad(f, Sigma_inv, Q, m, sigma),
                                  hess=lambda f: -O.T_hessian(f, Sigma_inv, Q, m, sigma), options=opts, callback=cb)
    return res, times


def rff_tight(PhiX, Q, m, sigma, w0, iters=50, tol=1e-13):
    """stationary point of S next to w0: Newton with the exact weight-space Hessian -I - Psi' a Psi"""
    w = np.asarray(w0, float).copy()
    F = PhiX.shape[0]
    D3 = O._rff_diffs(PhiX, Q, m).reshape(F, Q * m)
    for _ in range(iters):
        g = O.rff_S_grad(w, PhiX, Q, m, sigma)
        a = O.arrow_coeffs(PhiX.T @ w, Q, m, sigma).ravel()
        H = np.eye(F) + (D3 * a[None, :]) @ D3.T
        step = np.linalg.solve(H, g)
        w = w + step
        if np.abs(step).max() <= tol * max(1.0, np.abs(w).max()):
            break
    return w


def make_fixture(name):
    prob = synthetic.make_problem(name)
    X, theta, Q, m = prob["X"], prob["theta"], prob["Q"], prob["m"]
    N = X.shape[0]
    t0 = time.perf_counter()
    Sigma = O.regularize_covariance(O.se_kernel(X, X, theta), svd_roundtrip=False)
    Sinv = O.pd_inverse(Sigma)
    print("[%s] N=%d Sigma, Sigma^-1: %.1f s" % (name, N, time.perf_counter() - t0), flush=True)
    res, times = trust_exact_counted(Sinv, Q, m, theta[0], np.zeros(N))
    f_te = res.x
    print("[%s] trust-exact from 0: nit=%d %.1f s |grad|=%.2e" % (name, res.nit, sum(times), np.linalg.norm(res.jac)), flush=True)
    f_tight = O.fmap_tight(Sigma, Q, m, theta[0], f_te)
    g_tight = O.T_grad(f_tight, Sinv, Q, m, theta[0])
    print("[%s] tight: |f_te - f_tight|/|f| = %.2e  |grad T(f_tight)| = %.2e" % (
        name, np.abs(f_te - f_tight).max() / np.abs(f_tight).max(), np.linalg.norm(g_tight)), flush=True)
    alpha = scipy.linalg.cho_solve(scipy.linalg.cho_factor(Sigma, lower=True), f_tight)
    arrow = O.arrow_coeffs(f_tight, Q, m, theta[0]).ravel()
    out = dict(name=name, N=N, Q=Q, m=m, theta=np.array(theta), f_trust_exact=f_te, trust_exact_nit=res.nit,
               f_tight=f_tight, grad_norm_tight=float(np.linalg.norm(g_tight)), T_tight=float(O.T_value(f_tight, Sinv, Q, m, theta[0], quadrature=False)),
               n_neg=int((arrow < 0).sum()))
    # mu* over the bench pipeline's candidate set, posterior mean on the grids
    grids = prob["grids"]
    mu_grid = np.stack([O.se_kernel(X, g, theta).T @ alpha for g in grids])
    out["mu_grid"] = mu_grid
    out["mustar"] = float(max(f_tight.max(), mu_grid.max()))
    # exact-GP posterior covariance on three grids, 64 points each (subsampled: P x P at P = 1024 would be 8 MB per grid)
    sub = np.linspace(0, grids.shape[1] - 1, 64).astype(int)
    Lam = O.create_Lambda(f_tight, Q, m, theta[0])
    W = -Lam
    Bm = np.linalg.solve(np.eye(N) + W @ Sigma, W)            # W (I + Sigma W)^-1 == Sigma^-1 - Sigma^-1 P Sigma^-1, stable form
    covs = []
    for b in (0, grids.shape[0] // 2, grids.shape[0] - 1):
        g = grids[b][sub]
        k = O.se_kernel(X, g, theta)
        Kss = O.regularize_covariance(O.se_kernel(g, g, theta), svd_roundtrip=False)
        covs.append(Kss - k.T @ Bm @ k)
    out["cov_grid_ids"] = np.array([0, grids.shape[0] // 2, grids.shape[0] - 1])
    out["cov_sub"] = sub
    out["cov_grids"] = np.stack(covs)
    del Lam, W, Bm, Sinv
    # weight space
    Wf, bf, F = prob["W"], prob["b"], prob["F"]
    PhiX = O.rff_features(Wf, bf, X, theta[2])
    w_te, r2 = O.rff_omega_map(PhiX, Q, m, theta[0], np.zeros(F))
    w_tight = rff_tight(PhiX, Q, m, theta[0], w_te)
    hd = O.rff_S_hess_diag(w_tight, PhiX, Q, m, theta[0])
    print("[%s] rff: trust-exact nit=%d |w_te - w_tight|/|w| = %.2e |grad S| = %.2e" % (
        name, r2.nit, np.abs(w_te - w_tight).max() / np.abs(w_tight).max(), np.linalg.norm(O.rff_S_grad(w_tight, PhiX, Q, m, theta[0]))), flush=True)
    out.update(omega_tight=w_tight, hess_diag=hd, rff_trust_exact_nit=r2.nit)
    S = min(SLICE, prob["S"])
    Z = O.philox_normals(SEED, 0, 0, S * F).reshape(S, F)
    Omega = O.rff_sample_omega(w_tight, hd, Z)
    B = grids.shape[0]
    fmax = np.empty((B, S))
    arg = np.empty((B, S), dtype=np.int16)
    gap = np.empty((B, S), dtype=np.float32)
    for b in range(B):
        Fs = Omega @ O.rff_features(Wf, bf, grids[b], theta[2])
        idx = Fs.argmax(axis=1)
        top = Fs[np.arange(S), idx]
        Fs[np.arange(S), idx] = -np.inf
        fmax[b], arg[b], gap[b] = top, idx, top - Fs.max(axis=1)
    sums = np.stack([np.maximum(fmax - out["mustar"], 0).sum(axis=1), fmax.sum(axis=1), (fmax ** 2).sum(axis=1)], axis=1)
    out.update(slice_samples=S, seed=SEED, slice_fmax=fmax, slice_arg=arg, slice_gap=gap, slice_sums=sums,
               slice_direction=int(np.argmax(sums[:, 0])))
    path = os.path.join(ROOT, "tests", "golden", "full_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("[%s] -> %s (%.0f KB) direction %d mustar %.6g n_neg %d" % (name, path, os.path.getsize(path) / 1024, out["slice_direction"],
                                                                      out["mustar"], out["n_neg"]), flush=True)


def make_record(name, sizes=None):
    """Cost record of the reference algorithm: full trust-exact run from the reference's random start at full N (iteration count
    and seconds per outer iteration), the same at smaller N (scaling check), and the other N^3 pieces."""
    cfg = synthetic.CONFIGS[name]
    m = cfg["m"]
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    rec = {"problem": name, "host_threads": threads, "cpu_count": os.cpu_count(), "runs": []}
    for Qs in (sizes or [cfg["Q"] // 4, cfg["Q"] // 2, cfg["Q"]]):
        prob = synthetic.make_problem(name, Q=Qs)
        X, theta = prob["X"], prob["theta"]
        N = X.shape[0]
        r = {"Q": Qs, "N": N}
        t = time.perf_counter()
        Sigma = O.regularize_covariance(O.se_kernel(X, X, theta), svd_roundtrip=False)
        r["gram_s"] = time.perf_counter() - t
        t = time.perf_counter()
        Sinv = O.pd_inverse(Sigma)
        r["Sigma_inverse_s"] = time.perf_counter() - t
        rng = np.random.RandomState(0)
        f0 = np.linalg.cholesky(Sigma) @ rng.standard_normal(N)      # f ~ N(0, Sigma) (the reference draws it through an SVD factor)
        res, times = trust_exact_counted(Sinv, Qs, m, theta[0], f0)
        r.update(trust_exact_nit=int(res.nit), trust_exact_nfev=int(res.nfev), trust_exact_nhev=int(res.nhev),
                 trust_exact_total_s=float(sum(times)), trust_exact_s_per_iteration=float(np.mean(times)),
                 trust_exact_s_per_iteration_median=float(np.median(times)), grad_norm=float(np.linalg.norm(res.jac)))
        t = time.perf_counter()
        O.posterior_covariance(Sinv, res.x, Qs, m, theta[0])
        r["posterior_covariance_s"] = time.perf_counter() - t
        print(json.dumps(r), flush=True)
        rec["runs"].append(r)
    path = os.path.join(ROOT, "profiles", "r02_reference_full_fit_%s.json" % name)
    with open(path, "w") as fh:
        json.dump(rec, fh, indent=1)
    print("->", path)


if __name__ == "__main__":
    mode, name = sys.argv[1], sys.argv[2]
    if mode == "fixture":
        make_fixture(name)
    else:
        make_record(name, [int(s) for s in sys.argv[3:]] or None)
