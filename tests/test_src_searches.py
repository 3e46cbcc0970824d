"""GPU tests of the reference-API entry points that wrap an outer search around the device path:
  * maximize_EI / maximize_varmax / maximize_EI_fixed_x / maximize_varmax_given_xi (src/acquisition.py:91-131,189-218) through
    next_query's 'EI' / 'EXR' / 'EI-FIXEDX' / 'COORDINATE-VARMAX' strategies.  GPyOpt is absent, so the searches score a uniform
    candidate set of the same budget in one batched device pass; there is no reference trajectory to replay.
    What is pinned: the query structure the reference produces (which coordinates of xi / x are free, zero, perturbed,
    normalised), the cyclic state, and that the returned query IS the best-scoring candidate of the replayed RNG stream;
  * Hsampler.sample_xstar / sample_xstar_for_dim / return_xstar_for_dim (src/random_fourier_sampler.py:143-227): feasible, and no
    worse than every start point the replayed RNG stream hands the local optimiser (L-BFGS-B and Nelder-Mead never return a
    point worse than their start);
  * GPModel.optimize_theta and update_model(optimize_theta=True) (src/gp_model.py:354-366,391-413): reference bounds, sigma pinned
    to 1, a finite evidence at the optimum that is the best of the replayed search, Sigma rebuilt at the new theta."""
import numpy as np
import pytest

from conftest import relerr
from test_src_gpu import _model

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _normalised(xi):
    return np.abs(xi) / np.max(np.abs(xi))


@pytest.mark.parametrize("strategy", ["EI", "EXR", "EI-FIXEDX"])
def test_outer_search_over_xi_and_x(golden, strategy):
    import acquisition
    g = golden
    D = g["D"]
    st, gp = _model(g, strategy=strategy)
    st.mc_samples, st.BO_maxiter = 40, 11                    # 5 + 11 = 16 candidates of 40 posterior draws each
    prev = list(st.xi_dims_prev_iter)
    xi_dims = [int(d) for d in (np.array(prev) + 1) % D]      # the cyclic rule of next_query (:13-15)
    x_dims = [i for i in range(D) if i not in xi_dims]
    np.random.seed(11)
    xi, x = acquisition.next_query(st, gp, unscale=False)
    assert [int(d) for d in st.xi_dims_prev_iter] == xi_dims
    assert xi.shape == (D,) and x.shape == (D,)
    assert np.all(xi[x_dims] == 0) and np.all(xi[xi_dims] > 0) and xi.max() == 1.0
    assert np.all(x[xi_dims] == 0) and np.all((x[x_dims] > 0) & (x[x_dims] <= 1))
    if strategy == "EI-FIXEDX":
        assert np.array_equal(x[x_dims], acquisition.perturbate_zerocoordinates(gp.xstar.copy(), x_dims)[x_dims])
    # replay: same seed -> same candidates -> same posterior draws per candidate -> same scores; the query is the best of them
    values = acquisition._varmax_values if strategy == "EXR" else acquisition._ei_values
    np.random.seed(11)
    if strategy == "EI-FIXEDX":
        cand = np.random.uniform(0, 1, (5 + st.BO_maxiter, len(xi_dims)))

        def pair(v):
            xi_ = gp.xstar.copy()
            xi_[xi_dims] = v
            return (xi_, gp.xstar.copy())
        pairs = [pair(c) for c in cand]
    else:
        cand = np.random.uniform(0, 1, (5 + st.BO_maxiter, D))
        pairs = [acquisition._split(c, xi_dims, x_dims, D) for c in cand]
    scores = values(pairs, gp, st.mc_samples)
    assert scores.shape == (len(cand),) and np.all(np.isfinite(scores)) and np.all(scores >= 0)
    best = cand[int(np.argmax(scores))]
    xi_best = np.zeros(D)
    xi_best[xi_dims] = best if strategy == "EI-FIXEDX" else best[xi_dims]
    xi_best = _normalised(acquisition.perturbate_zerocoordinates(xi_best, xi_dims))
    assert np.array_equal(xi, xi_best)
    if strategy != "EI-FIXEDX":
        x_best = np.zeros(D)
        x_best[x_dims] = best[x_dims]
        assert np.array_equal(x, acquisition.perturbate_zerocoordinates(x_best, x_dims))
    # the single-pair callbacks GPyOpt would drive score one candidate like the batch does (same draws: re-seed per call)
    one = acquisition.varmax_to_maximize if strategy == "EXR" else acquisition.EI_to_maximize
    if strategy != "EI-FIXEDX":
        np.random.seed(5)
        a = float(one(cand[:1], xi_dims, x_dims, gp, st.mc_samples))
        np.random.seed(5)
        b = float(values(pairs[:1], gp, st.mc_samples)[0])
        assert a == b


def test_varmax_x_rule(golden):
    """x_acquisition_function 'varmax' (COORDINATE-VARMAX: xi = e_d by the cyclic rule, x from maximize_varmax_given_xi)"""
    import acquisition
    g = golden
    D = g["D"]
    st, gp = _model(g, strategy="COORDINATE-VARMAX")
    assert st.x_acquisition_function == "varmax"
    st.mc_samples, st.BO_maxiter = 40, 11
    np.random.seed(12)
    xi, x = acquisition.next_query(st, gp, unscale=False)
    d = int(st.dim_query_prev_iter) - 1
    assert d == 0 and np.array_equal(xi, np.eye(D)[0])        # the cycle starts at the first coordinate
    free = [i for i in range(D) if i != d]
    assert x[d] == 0 and np.all((x[free] > 0) & (x[free] <= 1))
    np.random.seed(12)
    cand = np.random.uniform(0, 1, (5 + st.BO_maxiter, D))
    scores = acquisition._varmax_values([(np.eye(D)[0], c) for c in cand], gp, st.mc_samples)
    best = cand[int(np.argmax(scores))].copy()
    best[d] = 0
    assert np.array_equal(x, acquisition.perturbate_zerocoordinates(best, free))
    # second call advances the cycle
    xi2, _ = acquisition.next_query(st, gp, unscale=False)
    assert np.array_equal(xi2, np.eye(D)[1 % D])


def _sampler(g, gp):
    from random_fourier_sampler import Hsampler
    from ppbo_b200 import ops
    F = g["rff_W"].shape[0]
    np.random.seed(int(g["seed_rff"]))
    h = Hsampler(gp, nFeatures=F)
    h.generate_basis()
    h.update_phi_X()
    h.update_omega_MAP(omega_initial=g["rff_omega_MAP"])
    h.update_covariancematrix()
    return h


def test_sample_xstar(golden):
    """sample_xstar = return_xstar(sample_omega()) (src/random_fourier_sampler.py:143-172,215-220): the RNG stream is one posterior
    draw of omega, then per restart one start index and D jitters.  L-BFGS-B is monotone, so the returned maximiser is at least
    as good as each of the five starts."""
    g = golden
    if "rff_W" not in g:
        pytest.skip("RFF basis exists for the SE kernel only")
    st, gp = _model(g)
    h = _sampler(g, gp)
    D = g["D"]
    state = np.random.RandomState(21).get_state()
    np.random.set_state(state)
    xs = h.sample_xstar()
    assert xs is not None and xs.shape == (D,) and np.all((xs >= 0) & (xs <= 1))
    np.random.set_state(state)
    omega = h.sample_omega()
    loc = h.GP_xstars_local
    starts = []
    for _ in range(5):
        x0 = loc[np.random.randint(loc.shape[0])]
        starts.append(np.clip(x0 + 0.01 * np.random.uniform(0, 1, size=D), 0, 1))
    val = float(h.phi(xs) @ omega)
    start_vals = np.array([float(h.phi(s) @ omega) for s in starts])
    scale = max(np.abs(start_vals).max(), abs(val), 1e-300)
    assert val >= start_vals.max() - 1e-12 * scale, (val, start_vals)
    # device value / gradient of the sampled function behind the search, against the host features
    from ppbo_b200 import ops
    v, gr = h._value_grad(ops.to_dev(omega), xs)
    assert abs(v - val) <= 1e-12 * scale
    gr_host = h.Dphi(xs).T @ omega                            # Dphi is F x D
    assert np.abs(gr - gr_host).max() <= 1e-10 * max(np.abs(gr_host).max(), scale)


def test_sample_xstar_for_dim(golden):
    """coordinate-wise maximiser (src/random_fourier_sampler.py:174-204,222-227): only coordinate `dim` of x_ref moves, it stays in
    [0, 1], and Nelder-Mead's best vertex is no worse than its start (x_ref with the GP maximiser's coordinate)"""
    g = golden
    if "rff_W" not in g:
        pytest.skip("RFF basis exists for the SE kernel only")
    st, gp = _model(g)
    h = _sampler(g, gp)
    D = g["D"]
    dim = min(2, D)
    x_ref0 = np.clip(g["xstar"] + 0.05, 0, 1)
    state = np.random.RandomState(22).get_state()
    np.random.set_state(state)
    omega = h.sample_omega()
    np.random.set_state(state)
    xs = h.sample_xstar_for_dim(dim, x_ref0.copy())
    assert xs.shape == (D,) and 0 <= xs[dim - 1] <= 1
    keep = [i for i in range(D) if i != dim - 1]
    assert np.array_equal(xs[keep], x_ref0[keep])
    start = x_ref0.copy()
    start[dim - 1] = h.GP_xstar[dim - 1]
    v_start, v_end = float(h.phi(start) @ omega), float(h.phi(xs) @ omega)
    assert v_end >= v_start - 1e-12 * max(abs(v_start), abs(v_end), 1e-300)


def test_optimize_theta(golden):
    """the evidence search (src/gp_model.py:391-413) without GPyOpt: scipy differential evolution over the reference's box with
    sigma = 1; the returned theta is the best point of the replayed search and update_model rebuilds Sigma with it"""
    import scipy.optimize
    from ppbo_b200 import ops
    g = golden
    st, gp = _model(g)
    np.random.seed(31)
    gp.optimize_theta()
    th = [float(t) for t in gp.theta]
    assert th[0] == 1.0 and 0.01 <= th[1] <= 2 and 0.1 <= th[2] <= 15
    np.random.seed(32)
    ev = gp.evidence(th, None)
    assert ev != -500 and np.isfinite(ev)
    # replay of the search on the same RNG stream with every evaluation recorded
    seen = []

    def neg(t):
        v = gp.evidence([1.0, t[0], t[1]], None)
        seen.append((v, float(t[0]), float(t[1])))
        return -v
    np.random.seed(31)
    res = scipy.optimize.differential_evolution(neg, [(0.01, 2), (0.1, 15)], maxiter=5, popsize=5, polish=False)
    assert [1.0, float(res.x[0]), float(res.x[1])] == th
    vals = np.array([s[0] for s in seen])
    assert len(seen) >= 20 and abs(max(vals) + res.fun) <= 1e-9 * abs(res.fun)
    assert np.sum(vals > -500) >= len(vals) // 2              # the sentinel is the exception, not the landscape
    # update_model(optimize_theta=True): theta replaced, Sigma rebuilt with it
    st2, gp2 = _model(g, fit=False)
    gp2.mustar_method = "batched"
    np.random.seed(33)
    gp2.update_model(optimize_theta=True)
    th2 = [float(t) for t in gp2.theta]
    assert th2[0] == 1.0 and 0.01 <= th2[1] <= 2 and 0.1 <= th2[2] <= 15
    Sig = ops.gram_regularized(gp2._kernel_name(), gp2._Xd(), th2[1], th2[2], gp2.COVARIANCE_SHRINKAGE).cpu().numpy()
    assert np.array_equal(np.asarray(gp2.Sigma), Sig)
    assert gp2.xstar.shape == (g["D"],) and np.isfinite(gp2.mustar)
