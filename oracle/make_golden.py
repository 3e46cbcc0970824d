"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING the unmodified reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py            # writes tests/golden/<case>.npz

Every array stored is either an input handed to the reference or an output the reference's own
code produced (GPModel, acquisition.EI/varmax/next_query, Hsampler).  All random draws come from
the legacy global numpy RNG, seeded immediately before each reference call; the seeds are stored
so the oracle / CUDA path can replay the identical draws in the identical order.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

warnings.filterwarnings("ignore")

CASES = {
    # name: D, bounds, Q, m, kernel, theta        (shapes follow SURVEY.md 8, shrunk to run in seconds)
    "camel2d": dict(D=2, bounds=((-3, 3), (-2, 2)), Q=6, m=25, kernel="SE_kernel", theta=[0.01, 0.26, 0.1]),
    "hartmann6d": dict(D=6, bounds=((0, 1),) * 6, Q=8, m=25, kernel="SE_kernel", theta=[0.001, 0.26, 0.1]),
    "camphor6d": dict(D=6, bounds=((-.5, .5), (-.5, .5), (4, 7), (-180, 180), (-180, 180), (-180, 180)), Q=6, m=25,
                      kernel="camphor_copper_kernel", theta=[0.001, 0.26, 0.1]),
    "levy10d": dict(D=10, bounds=((-10, 10),) * 10, Q=10, m=12, kernel="SE_kernel", theta=[0.001, 0.4, 0.15]),
    "ackley20d": dict(D=20, bounds=((-32.768, 32.768),) * 20, Q=7, m=25, kernel="SE_kernel", theta=[0.09, 0.3, 0.5]),
    "rq3d": dict(D=3, bounds=((-1, 1), (0, 2), (-5, 5)), Q=5, m=10, kernel="RQ_kernel", theta=[0.05, 0.3, 0.2]),
}
# BASELINE configs 1-3 at the sizes the reference's own drivers reach (ppbo_numerical_main.py:131-183,186 and notebook cells
# 10-15: 4+35 / 6+35 / 6+14 queries, m = 25 -> N = 1014 / 1066 / 520).  Minutes of reference CPU each; the N x N matrices are
# not stored (8 MB each) -- SAMPLE_ROWS rows of each are.
CASES_FULL = {
    "camel2d_full": dict(CASES["camel2d"], Q=39),
    "hartmann6d_full": dict(CASES["hartmann6d"], Q=41),
    "camphor6d_full": dict(CASES["camphor6d"], Q=20),
}
SAMPLE_ROWS = 12
SEED_DATA, SEED_FIT, SEED_PRED, SEED_EI, SEED_VARMAX, SEED_QUERY, SEED_RFF = 11, 12, 13, 14, 15, 16, 17
SEED_XSTAR, SEED_EVIDENCE, SEED_EXTRA = 18, 19, 20
RFF_FEATURES, RFF_SAMPLES, RFF_GRID = 96, 24, 40
XSTAR_SAMPLES = 6
# (l, sigma_f) pairs relative to the case's own theta at which GPModel.evidence is recorded (sigma kept)
EVIDENCE_SCALES = ((1.0, 1.0), (1.5, 1.0), (0.8, 1.3))


def synthetic_queries(ref, D, bounds, Q, utility_centre):
    """Query log rows [alpha*xi + x | xi | alpha*] in original units (ppbo_numerical_main.py:63-69).
    The 'user' maximises a smooth synthetic utility along the line, so the data are self-consistent."""
    lo = np.array([b[0] for b in bounds], dtype=float)
    hi = np.array([b[1] for b in bounds], dtype=float)
    rows = []
    for q in range(Q):
        xi = np.zeros(D)
        if q < D or q % 3:
            xi[q % D] = hi[q % D] if hi[q % D] != 0 else 1.0
        else:                                   # a few genuinely projective (multi-coordinate) queries
            # (only over coordinates whose range contains 0: alpha xi must be able to stay inside the box together with x = 0
            # there -- the camphor problem's height coordinate (4, 7) cannot be part of a multi-coordinate direction)
            ok = np.where((lo <= 0) & (hi >= 0))[0]
            dims = np.random.choice(D if len(ok) == D else ok, size=min(len(ok), 2), replace=False)
            xi[dims] = np.random.uniform(0.2, 1.0, size=len(dims)) * np.abs(hi[dims])
        x = np.random.uniform(lo, hi)
        x[xi != 0] = 0
        amin, amax = ref.misc.alpha_bounds(xi, lo, hi)
        alphas = np.linspace(amin, amax, 401)
        pts = alphas[:, None] * xi[None, :] + x[None, :]
        util = -np.sum(((pts - utility_centre) / (hi - lo)) ** 2, axis=1)
        a_star = float(alphas[np.argmax(util)]) + 1e-3 * (amax - amin) * np.random.randn()
        a_star = float(np.clip(a_star, amin, amax))
        rows.append(np.concatenate([a_star * xi + x, xi, [a_star]]))
    return np.array(rows)


def run_case(name, cfg, full=False):
    ref = ref_shim.load()
    import scipy.optimize
    D, bounds, Q, m = cfg["D"], cfg["bounds"], cfg["Q"], cfg["m"]
    out = dict(D=D, Q=Q, m=m, bounds=np.array(bounds, dtype=float), kernel=cfg["kernel"],
               theta=np.array(cfg["theta"], dtype=float))
    np.random.seed(SEED_DATA)
    lo = np.array([b[0] for b in bounds], dtype=float)
    hi = np.array([b[1] for b in bounds], dtype=float)
    centre = lo + np.random.uniform(0.25, 0.75, D) * (hi - lo)
    X_obs = synthetic_queries(ref, D, bounds, Q, centre)
    out["X_obs"] = X_obs

    settings = ref.ppbo_settings.PPBO_settings(D=D, bounds=bounds, xi_acquisition_function="EI-EXT-FAST", m=m,
                                               theta_initial=list(cfg["theta"]), kernel=cfg["kernel"], verbose=False,
                                               alpha_grid_distribution="equispaced")
    gp = ref.gp_model.GPModel(settings)
    out["seed_design"] = SEED_DATA + 100
    np.random.seed(SEED_DATA + 100)
    gp.update_feedback_processing_object(X_obs)
    gp.update_data()
    gp.turn_initialization_off()
    out["X"] = gp.X.copy()
    out["X_full"] = gp.FP.X_full.copy()
    out["obs_indices"] = np.array(gp.obs_indices)

    # ---- fit (capture the start vector handed to scipy)
    captured = {}
    real_minimize = scipy.optimize.minimize

    def spy(fun, x0, *a, **k):
        captured.setdefault("x0", np.array(x0, dtype=float).ravel().copy())
        res = real_minimize(fun, x0, *a, **k)
        captured.setdefault("nit", res.nit)
        captured.setdefault("xmin", np.array(res.x, dtype=float).ravel().copy())
        return res
    scipy.optimize.minimize = spy
    try:
        np.random.seed(SEED_FIT)
        gp.update_model()
    finally:
        scipy.optimize.minimize = real_minimize
    theta = gp.theta
    out.update(seed_fit=SEED_FIT, f_initial=captured["x0"], fit_nit=captured["nit"], fMAP=gp.fMAP.copy(),
               xstar=np.array(gp.xstar, dtype=float).reshape(D), mustar=float(gp.mustar),
               xstars_local=np.array(gp.xstars_local, dtype=float).reshape(-1, D))
    if full:
        rows = np.linspace(0, gp.N - 1, SAMPLE_ROWS).astype(int)
        out.update(sample_rows=rows, K_raw_rows=gp.kernel(gp.X[rows], gp.X, theta), Sigma_rows=gp.Sigma[rows].copy(),
                   Sigma_inv_rows=gp.Sigma_inv[rows].copy(), posterior_covariance_rows=gp.posterior_covariance[rows].copy(),
                   Sigma_diag=np.diag(gp.Sigma).copy(), posterior_covariance_diag=np.diag(gp.posterior_covariance).copy())
    else:
        out.update(K_raw=gp.kernel(gp.X, gp.X, theta), Sigma=gp.Sigma.copy(), Sigma_inv=gp.Sigma_inv.copy(),
                   posterior_covariance=gp.posterior_covariance.copy())
    Lam = gp.Lambda_MAP
    out["Lambda_MAP_rows"] = np.array([Lam[i, i:i + m + 1] for i in gp.obs_indices])   # arrow rows (winner row)
    out["Lambda_MAP_diag"] = np.diag(Lam).copy()
    out["Lambda_MAP_nnz"] = int(np.count_nonzero(Lam))
    for tag, f in (("init", captured["x0"]), ("map", gp.fMAP)):
        out["T_" + tag] = float(gp.T(f, theta))
        out["T_grad_" + tag] = gp.T_grad(f, theta)
    Lam0 = gp.create_Lambda(captured["x0"], theta[0])
    out["Lambda_init_rows"] = np.array([Lam0[i, i:i + m + 1] for i in gp.obs_indices])
    out["Lambda_init_diag"] = np.diag(Lam0).copy()

    # ---- prediction on a projected grid + a rectangular kernel block
    np.random.seed(SEED_PRED)
    xi = np.zeros(D)
    xi[D // 2] = 1.0
    xq = gp.xstar.copy()
    xq[D // 2] = 0.0
    grid = gp.FP.xi_grid(xi=xi, x=xq, alpha_grid_distribution="equispaced", alpha_star=None, m=70, is_scaled=True)
    mu, Sp = gp.mu_Sigma_pred(grid)
    out.update(seed_pred=SEED_PRED, pred_grid=grid, pred_mu=mu, pred_Sigma=Sp, mu_pred_xstar=gp.mu_pred(gp.xstar))
    if full:
        out["K_cross_rows"] = gp.kernel(gp.X[out["sample_rows"]], grid, theta)
    else:
        out["K_cross"] = gp.kernel(gp.X, grid, theta)

    # ---- acquisition: per-direction EI / varmax exactly as EId_xstar walks them (src/acquisition.py:132-145)
    S = settings.mc_samples
    for tag, seed, fn in (("EI", SEED_EI, ref.acquisition.EI), ("varmax", SEED_VARMAX, ref.acquisition.varmax)):
        np.random.seed(seed)
        vals = []
        for d in range(D):
            e = np.zeros(D)
            e[d] = 1.0
            xs = gp.xstar.copy()
            xs[d] = 0
            vals.append(fn(e, xs, gp, S))
        out[tag + "_vals"] = np.array(vals)
        out["seed_" + tag] = seed
    out["mc_samples"] = S
    np.random.seed(SEED_QUERY)
    xi_n, x_n = ref.acquisition.next_query(settings, gp, unscale=True)
    out.update(seed_query=SEED_QUERY, next_xi=xi_n, next_x=x_n)
    for strat in ("PCD", "EXT"):
        st = ref.ppbo_settings.PPBO_settings(D=D, bounds=bounds, xi_acquisition_function=strat, m=m,
                                             theta_initial=list(cfg["theta"]), kernel=cfg["kernel"], verbose=False)
        qs = []
        for _ in range(3):
            a, b = ref.acquisition.next_query(st, gp, unscale=True)
            qs.append(np.concatenate([a, b]))
        out["next_" + strat] = np.array(qs)

    # ---- random Fourier features (SE kernel only, src/random_fourier_sampler.py:40-42)
    if cfg["kernel"] == "SE_kernel":
        np.random.seed(SEED_RFF)
        h = ref.random_fourier_sampler.Hsampler(gp, nFeatures=RFF_FEATURES)
        h.generate_basis()
        h.update_phi_X()
        omega_probe = np.random.randn(RFF_FEATURES)
        if full:
            out["rff_phi_X_cols"] = h.phi_X[:, out["sample_rows"]].copy()
        else:
            out["rff_phi_X"] = h.phi_X.copy()
        out.update(seed_rff=SEED_RFF, rff_W=h.W.copy(), rff_b=h.b.ravel().copy(),
                   rff_omega_probe=omega_probe, rff_S_probe=float(h.S(omega_probe, theta)),
                   rff_S_grad_probe=h.S_grad(omega_probe, theta),
                   rff_S_hess_diag_probe=np.diag(h.S_hessian(omega_probe, theta)).copy(),
                   rff_Dphi_xstar=h.Dphi(gp.xstar))
        captured.clear()
        scipy.optimize.minimize = spy
        try:
            h.update_omega_MAP()
        finally:
            scipy.optimize.minimize = real_minimize
        h.update_covariancematrix()
        out.update(rff_omega0=captured["x0"], rff_omega_MAP=h.omega_MAP.copy(),
                   rff_cov_diag=np.diag(h.covariance).copy(),
                   rff_cov_offdiag_max=float(np.abs(h.covariance - np.diag(np.diag(h.covariance))).max()))
        Omega = np.array([h.sample_omega() for _ in range(RFF_SAMPLES)])
        gridr = gp.FP.xi_grid(xi=xi, x=xq, alpha_grid_distribution="equispaced", alpha_star=None, m=RFF_GRID,
                              is_scaled=True)
        # the reference evaluates a sampled function one point at a time (src/random_fourier_sampler.py:166,170)
        Fs = np.array([[float(np.dot(h.phi(x).T, om)) for x in gridr] for om in Omega])
        out.update(rff_Omega=Omega, rff_grid=gridr, rff_Fs=Fs, rff_max=Fs.max(axis=1),
                   rff_argmax=Fs.argmax(axis=1).astype(np.int32))
        # ---- maximiser of sampled functions (Hsampler.return_xstar, src/random_fourier_sampler.py:143-178)
        np.random.seed(SEED_XSTAR)
        xs_, vals_ = [], []
        for om in Omega[:XSTAR_SAMPLES]:
            xs = h.return_xstar(om)
            xs_.append(xs)
            vals_.append(float(np.dot(h.phi(xs).T, om)))
        out.update(seed_xstar=SEED_XSTAR, rff_xstar=np.array(xs_), rff_xstar_val=np.array(vals_))

    # ---- more acquisition entry points and the dense Hessian (recorded after everything above so that the earlier arrays keep
    # their values): EId_integrate (src/acquisition.py:146-163), T_hessian rows (src/gp_model.py:242-247), Hsampler.sum_Phi
    np.random.seed(SEED_EXTRA)
    if not full:
        out["EId_integrate_xi"] = ref.acquisition.EId_integrate(gp, 20)
        out["seed_extra"] = SEED_EXTRA
        H = gp.T_hessian(gp.fMAP, theta)
        out["T_hessian_map_rows"] = H[np.array(gp.obs_indices)[:3]].copy()
        out["sum_Phi_vec_map"] = np.array([gp.sum_Phi_vec(o, gp.fMAP, theta[0]) for o in (0, 1, 2)], dtype=float).reshape(3, -1)
        if cfg["kernel"] == "SE_kernel":
            fw = np.dot(h.phi_X.T, omega_probe)
            i0 = int(gp.obs_indices[1])
            sp, wq = np.polynomial.hermite.hermgauss(h.n_gausshermite_sample_points)
            out["rff_sum_Phi_probe0"] = float(h.sum_Phi(i0, 0, fw, theta[0], sp, wq))
            out["rff_sum_Phi_probe1"] = np.asarray(h.sum_Phi(i0, 1, fw, theta[0]), dtype=float).ravel()
            out["rff_sum_Phi_probe2"] = np.asarray(h.sum_Phi(i0, 2, fw, theta[0]), dtype=float).ravel()

    # ---- evidence (src/gp_model.py:278-319): value, the mode its own trust-exact run found, and the LU log-determinant pieces
    if not full:
        import scipy.linalg
        ev_t, ev_v, ev_f, ev_sign, ev_logabs, ev_T = [], [], [], [], [], []
        for k, (sl, sf) in enumerate(EVIDENCE_SCALES):
            th = [theta[0], theta[1] * sl, theta[2] * sf]
            captured.clear()
            scipy.optimize.minimize = spy
            try:
                np.random.seed(SEED_EVIDENCE + k)
                val = gp.evidence(th, None)
            finally:
                scipy.optimize.minimize = real_minimize
            # the minimiser returns the mode the evidence was evaluated at: re-run the capture through res.x
            ev_t.append(th)
            ev_v.append(float(val))
            ev_f.append(captured["xmin"])
            Sig = gp.create_Gramian(gp.X, gp.X, gp.kernel, th)
            Mx = np.eye(gp.N) + Sig.dot(gp.create_Lambda(captured["xmin"], th[0]))
            sg, la = np.linalg.slogdet(Mx)
            ev_sign.append(sg)
            ev_logabs.append(la)
            ev_T.append(float(gp.T(captured["xmin"], th, ref.misc.pd_inverse(Sig))))
        out.update(seed_evidence=SEED_EVIDENCE, evidence_theta=np.array(ev_t), evidence_value=np.array(ev_v),
                   evidence_fmap=np.array(ev_f), evidence_det_sign=np.array(ev_sign), evidence_logabsdet=np.array(ev_logabs),
                   evidence_T=np.array(ev_T))
    return out


def main():
    if not ref_shim.available():
        raise SystemExit("reference not found; golden files can only be generated in the build container")
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    only = sys.argv[1:]
    todo = [(n, c, False) for n, c in CASES.items() if (not only or n in only)]
    todo += [(n, c, True) for n, c in CASES_FULL.items() if n in only]          # full-size cases only on request (minutes each)
    for name, cfg, full in todo:
        out = run_case(name, cfg, full)
        path = os.path.join(dst, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-12s N=%4d fit_nit=%4d |grad T(fMAP)|=%.2e  -> %s (%.0f KB)" % (
            name, out["X"].shape[0], out["fit_nit"], np.linalg.norm(out["T_grad_map"]), path,
            os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
