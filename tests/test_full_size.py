"""GPU parity at the BASELINE sizes (run with -m gpu on the B200).

Configs 4-5 (Levy-10D N = 2080, Ackley-20D N = 5200: the problems bench.py times) against the oracle fixtures
tests/golden/full_<name>.npz, which oracle/make_full_fixtures.py computed on the host INDEPENDENTLY of the CUDA path:
the reference's scipy trust-exact from f = 0 tightened by the oracle's own Newton iteration (never started from a CUDA result),
posterior mean / covariance with the stable W (I + Sigma W)^-1 form, the weight-space mode, and the FP64 sampled acquisition on
the first 2048 samples of the Philox stream.  These sizes run the parts the small goldens never reach: the look-ahead Cholesky
with 17-40 block columns, the rank-128 strip updates, the 1024-block inverse solves, Aitken steps and the rank-r Woodbury
correction for negative curvature coefficients (130 of 5000 on the Ackley problem).

Tolerances (BASELINE.json north_star): mode / posterior mean / variance 1e-6 relative (variance relative to sigma_f^2: it is a
cancellation); arg-max indices identical wherever the FP64 runner-up gap exceeds the rounding level of the
contraction (explicit tie rule below).
"""
import numpy as np
import pytest

from conftest import load_oracle_full, oracle_full_names, relerr

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SLICES = 6


def _np(t):
    return t.detach().cpu().numpy()


class _Case:
    """one cold-start pipeline run per problem, shared by the tests below"""

    def __init__(self, name):
        from ppbo_b200 import iteration, ops, synthetic
        self.name, self.fx = name, load_oracle_full(name)
        self.prob = p = synthetic.make_problem(name)
        self.ops, self.it = ops, iteration
        self.X = ops.to_dev(p["X"])
        self.theta, self.Q, self.m = p["theta"], p["Q"], p["m"]
        self.gp = iteration.gp_fit(self.X, p["kernel"], p["theta"], p["Q"], p["m"], tol=1e-9)          # cold start: f = 0
        self.W, self.b = ops.to_dev(p["W"]), ops.to_dev(p["b"])
        self.rff = iteration.rff_fit(self.X, self.W, self.b, p["theta"], p["Q"], p["m"], tol=1e-9)    # cold start: omega = 0
        self.grids = ops.to_dev(p["grids"])


_CASES = {}


@pytest.fixture(params=oracle_full_names())
def case(request):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if request.param not in _CASES:
        _CASES[request.param] = _Case(request.param)
    return _CASES[request.param]


def test_fixture_is_independent_of_cuda(case):
    """the yard-stick was produced by scipy trust-exact from zero + the oracle's Newton: a stationary point of the reference's T"""
    fx = case.fx
    assert fx["grad_norm_tight"] < 1e-5 and fx["trust_exact_nit"] >= 3
    assert np.abs(fx["f_trust_exact"] - fx["f_tight"]).max() <= 1e-4 * np.abs(fx["f_tight"]).max()


def test_mode_cold_start(case):
    f, f_ref = _np(case.gp.f_map), case.fx["f_tight"]
    assert np.abs(f - f_ref).max() <= 1e-6 * np.abs(f_ref).max(), np.abs(f - f_ref).max() / np.abs(f_ref).max()
    st = case.gp.lap.stats
    assert st["converged"] == 1 and st["iterations"] <= 40
    # the mode factor was NOT part of the fit (lazy): only the Newton factorisations
    assert st["factor_state"] == 1 and st["factorizations"] <= 10


def test_mode_gradient_reference_formula(case):
    """|grad T| = |-Sigma^-1 f + beta(f)| (src/gp_model.py:228-240) on the host, Sigma^-1 f by a host Cholesky solve"""
    import scipy.linalg
    from oracle import ppbo_oracle as O
    p = case.prob
    f = _np(case.gp.f_map)
    Sigma = O.regularize_covariance(O.se_kernel(p["X"], p["X"], p["theta"]), svd_roundtrip=False)
    assert relerr(_np(case.gp.Sigma), Sigma) < 1e-10
    a = scipy.linalg.cho_solve(scipy.linalg.cho_factor(Sigma, lower=True), f)
    gn = np.linalg.norm(-a + O.lik_beta(f, p["Q"], p["m"], p["theta"][0]))
    assert gn <= max(10 * float(case.fx["grad_norm_tight"]), 1e-5), gn
    assert relerr(_np(case.gp.alpha), a) < 1e-5        # alpha = Sigma^-1 f (cond(Sigma) ~ 1e6: relative 1e-5 of the largest entry)


def test_posterior_mean_all_grids_and_mustar(case):
    p, fx = case.prob, case.fx
    B, P, D = p["grids"].shape
    mu = _np(case.it.posterior_mean(case.gp, case.grids.reshape(B * P, D))).reshape(B, P)
    assert np.abs(mu - fx["mu_grid"]).max() <= 1e-6 * np.abs(fx["mu_grid"]).max()
    mustar = float(_np(case.it.mustar_over_candidates(case.gp, case.grids.reshape(B * P, D)))[0])
    assert abs(mustar - float(fx["mustar"])) <= 1e-6 * abs(float(fx["mustar"]))


def test_posterior_covariance_three_grids(case):
    """Sigma_p on three grids (64 points each) against the oracle's K** - k' W (I + Sigma W)^-1 k: runs the lazily built mode factor,
    the blocked right-TRSM and the Woodbury correction for the negative coefficients at full size"""
    p, fx, ops = case.prob, case.fx, case.ops
    ids, sub = fx["cov_grid_ids"], fx["cov_sub"]
    Xp = ops.to_dev(np.concatenate([p["grids"][b][sub] for b in ids]))
    lap = case.gp.lap
    mu, Sp = ops.predict(p["kernel"], case.X, p["theta"][1], p["theta"][2], 1e-6, lap, Xp, len(sub), len(ids))
    assert lap.factor_state == 2                       # built on demand
    assert lap.n_neg == int(fx["n_neg"])
    sf2 = p["theta"][2] ** 2
    err = np.abs(_np(Sp) - fx["cov_grids"]).max()
    assert err <= 1e-6 * sf2, err / sf2
    assert np.abs(_np(mu) - np.stack([fx["mu_grid"][b][sub] for b in ids])).max() <= 1e-6 * np.abs(fx["mu_grid"]).max()


def test_weight_space_mode(case):
    fx = case.fx
    w, hd = _np(case.rff.omega_map), _np(case.rff.hess_diag)
    assert np.abs(w - fx["omega_tight"]).max() <= 1e-6 * np.abs(fx["omega_tight"]).max()
    assert np.abs(hd - fx["hess_diag"]).max() <= 1e-6 * np.abs(fx["hess_diag"]).max()


def _oracle_rff(case):
    r = case.it.RFFFit()
    r.W, r.b, r.sigma_f = case.W, case.b, float(case.theta[2])
    r.omega_map, r.hess_diag = case.ops.to_dev(case.fx["omega_tight"]), case.ops.to_dev(case.fx["hess_diag"])
    return r


@pytest.mark.parametrize("engine", ["i8", "f64"])
def test_sampling_engines_identical_argmax(case, engine):
    """(iv) both sampling engines, fed the ORACLE's weight-space mode, against the oracle's FP64 contraction on the first 2048
    Philox samples: values to the accuracy of an FP64 GEMM of depth 1000, arg-max IDENTICAL wherever the oracle's gap between
    the best and the second-best grid point exceeds that accuracy (tie rule: below it either index is a correct arg-max)."""
    fx, it, ops = case.fx, case.it, case.ops
    S = int(fx["slice_samples"])
    r = _oracle_rff(case)
    PhiT = it.rff_grid_features(case.W, case.b, case.theta[2], case.grids)
    B, P, F = PhiT.shape
    Omega = ops.rff_sample_omega(r.omega_map, r.hess_diag, S, seed=int(fx["seed"]), sample0=0)
    if engine == "i8":
        fmax, arg, _ = ops.rff_eval_argmax_i8(Omega, PhiT, slices=SLICES)
    else:
        fmax, arg, _ = ops.rff_eval_argmax(Omega, PhiT)
    fmax, arg = _np(fmax), _np(arg)
    scale = np.abs(fx["slice_fmax"]).max()
    tol = 2e-12 * scale * np.sqrt(F)                   # rounding level of the contraction (both engines and numpy's own GEMM)
    assert np.abs(fmax - fx["slice_fmax"]).max() <= tol
    decided = fx["slice_gap"] > 1e-11 * scale          # tie rule: gaps at the rounding level of the FP64 contraction are ties
    assert decided.mean() > 0.995
    assert np.array_equal(arg[decided], fx["slice_arg"].astype(np.int32)[decided])
    # undecided samples: the chosen point must still be a maximiser within the tolerance
    assert np.all(fmax[~decided] >= fx["slice_fmax"][~decided] - tol)


def test_acquisition_slice_sums_and_direction(case):
    """(iii) the pipeline's own fits (CUDA modes, CUDA mu*) on the 2048-sample slice: sums and the selected direction"""
    fx, it, ops = case.fx, case.it, case.ops
    S = int(fx["slice_samples"])
    B, P, D = case.prob["grids"].shape
    PhiT = it.SlicedGrids(it.rff_grid_features(case.W, case.b, case.theta[2], case.grids))
    mustar = it.mustar_over_candidates(case.gp, case.grids.reshape(B * P, D))
    sums, fmax, arg = it.rff_acquisition(case.rff, PhiT, S, mustar, seed=int(fx["seed"]), bounds=(0, S))
    sums = _np(sums)
    ref = fx["slice_sums"]
    # sums of S numbers of size |fmax|: 1e-6 relative to the column scale (the EI column is a sum of clipped differences)
    for c in range(3):
        assert np.abs(sums[:, c] - ref[:, c]).max() <= 1e-6 * max(np.abs(ref[:, c]).max(), S * 1e-3 * np.abs(fx["slice_fmax"]).max())
    ei, var = it.acquisition_values(sums, S)
    assert int(np.argmax(ei)) == int(fx["slice_direction"])
    ei_ref = ref[:, 0] / S
    assert np.abs(ei - ei_ref).max() <= 1e-6 * max(ei_ref.max(), 1e-300) + 1e-9 * np.abs(fx["slice_fmax"]).max()


def test_warm_append_reaches_the_same_mode(case):
    """incremental growth: cold fit on Q-2 comparison sets, then two appends (new rows of Sigma and G
    bit-identical to from-scratch, the previous factor carried with the new rows as a border, chord iteration from the previous
    mode) -> the full problem's mode.  The cold fit stops after ONE factorisation when its mixed chord phase converges, so the first
    append may still pay for a fresh factor; from then on the iteration has no O(N^3) step."""
    p, fx, it, ops = case.prob, case.fx, case.it, case.ops
    m, Q = p["m"], p["Q"]
    st = it.GPState(p["kernel"], p["theta"], p["D"], m, Q, case.X.device, tol=1e-9)
    n2, n1 = (Q - 2) * (m + 1), (Q - 1) * (m + 1)
    st.cold(case.X[:n2])
    st.append(case.X[n2:n1])
    first = dict(st.lap.stats)
    st.append(case.X[n1:])
    assert torch.equal(st.Sigma, case.gp.Sigma)                       # appended rows / columns bit-identical
    G_ref = ops.diffspace_gram(case.gp.Sigma, Q, m)
    assert torch.equal(st.lap.G[:Q * m, :Q * m], G_ref)
    f = _np(st.f_map)
    assert np.abs(f - fx["f_tight"]).max() <= 1e-6 * np.abs(fx["f_tight"]).max()
    s = st.lap.stats
    assert s["converged"] == 1 and first["converged"] == 1
    # no O(N^3) step in the steady state on the well-conditioned problem; the sharp-likelihood one (sigma = 1e-3: a cold fit needs
    # 6 factorisations and 7 step halvings) may fall back to a Newton step or two
    assert first["factorizations"] <= (1 if case.name == "ackley20d" else 3), first
    assert s["factorizations"] <= (0 if case.name == "ackley20d" else 2), dict(s)
    # prediction from the grown model (mode factor built on demand at the new size)
    ids, sub = fx["cov_grid_ids"], fx["cov_sub"]
    Xp = ops.to_dev(np.concatenate([p["grids"][b][sub] for b in ids]))
    _, Sp = ops.predict(p["kernel"], st.X, p["theta"][1], p["theta"][2], 1e-6, st.lap, Xp, len(sub), len(ids))
    assert np.abs(_np(Sp) - fx["cov_grids"]).max() <= 1e-6 * p["theta"][2] ** 2


def test_overlapped_pipeline_equals_sequential(case):
    """One GPU: run_iteration with the sampling contraction overlapped with the GP fit (two host threads, contraction on a
    persistent grid that leaves SMs to the fit) returns bit for bit the sums of the sequential order -- cold and after an appended
    comparison set -- and those sums select the fixture's direction on its 2048-sample slice."""
    p, fx, it, ops = case.prob, case.fx, case.it, case.ops
    m, Q, S = p["m"], p["Q"], int(fx["slice_samples"])
    B, P, D = p["grids"].shape
    if it.sampling_engine(S, P, p["F"]) != "i8":
        pytest.skip("contraction below the INT8 threshold: nothing to overlap")
    n1 = (Q - 1) * (m + 1)
    d_cold = {"X": case.X[:n1].contiguous(), "W": case.W, "b": case.b, "grids": case.grids}
    d_warm = {"block": case.X[n1:].contiguous(), "W": case.W, "b": case.b, "grids": case.grids}
    out = {}
    saved = it.OVERLAP_SAMPLING
    try:
        for mode in (True, False):
            it.OVERLAP_SAMPLING = mode
            st = it.IterationState(p["kernel"], p["theta"], D, m, Q, case.X.device, case.W, case.b)
            s0, _, _ = it.run_iteration(d_cold, p["kernel"], p["theta"], Q - 1, m, S, seed=int(fx["seed"]), state=st)
            s1, gp, rff = it.run_iteration(d_warm, p["kernel"], p["theta"], Q, m, S, seed=int(fx["seed"]), state=st)
            torch.cuda.synchronize()
            out[mode] = (s0.clone(), s1.clone(), gp.f_map.clone(), rff.omega_map.clone())
    finally:
        it.OVERLAP_SAMPLING = saved
    for a, b_ in zip(out[True], out[False]):
        assert torch.equal(a, b_)
    ei, _ = it.acquisition_values(_np(out[True][1]), S)
    assert int(np.argmax(ei)) == int(fx["slice_direction"])
