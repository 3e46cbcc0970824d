"""CPU tests of the host-side differential evolution (ppbo_b200/csrc/de.cu, ppbo_de_minimize): GPModel.mu_star's sequential search
(src/gp_model.py:415-437: scipy.optimize.differential_evolution(mu_pred_neq, bounds, updating='immediate', maxiter=2000)) with
the loop in C++.  The claim is a REPLAY, so the checker is scipy itself on the same objective and the same numpy stream: identical
bits of x and fun, identical generation / evaluation counts, and the global generator left in the identical state.  The entry is
pure host code, so this runs without a GPU; on the GPU the same loop drives ppbo_mu_pred_point (tests/test_src_gpu.py)."""
import numpy as np
import pytest
import scipy.optimize

from ppbo_b200 import ops
from ppbo_b200._lib import PPBOError


def sphere(x):
    return float(np.sum((x - 0.3) ** 2))


def rastrigin(x):
    return float(10 * len(x) + np.sum(x * x - 10 * np.cos(2 * np.pi * x)))


def camel(x):                                  # numerical_experiments/test_functions.py: six-hump camel
    return float((4 - 2.1 * x[0] ** 2 + x[0] ** 4 / 3) * x[0] ** 2 + x[0] * x[1] + (-4 + 4 * x[1] ** 2) * x[1] ** 2)


def two_bumps(x):                              # shaped like a negated posterior mean: smooth, two maxima of different height
    return float(-np.exp(-np.sum((x - 0.7) ** 2) / 0.02) - 0.5 * np.exp(-np.sum((x - 0.2) ** 2) / 0.1))


def flat(x):                                   # every energy equal: all trials accepted, converged after one generation
    return 1.0


CASES = [(sphere, [(0, 1)] * 6, 3), (rastrigin, [(-5.12, 5.12)] * 3, 5), (camel, [(-3, 3), (-2, 2)], 7), (two_bumps, [(0, 1)] * 10, 11),
         (two_bumps, [(0, 1)] * 2, 13), (sphere, [(0, 1)], 19), (flat, [(0, 1)] * 3, 23), (sphere, [(0, 1), (0.5, 0.5), (-1, 2)], 29)]


def _same_state(a, b):
    return a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2:] == b[2:]


@pytest.mark.parametrize("window", [2, 16, 32, 64])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_speculative_windows_replay_scipy(case, window):
    """evaluations issued `window` trials at a time (one device launch per window on the GPU): the retained trials, draws and
    comparisons are the sequential loop's, so every bit and the generator state still equal scipy's; fewer calls, some discarded"""
    f, bounds, seed = CASES[case]
    np.random.seed(seed)
    np.random.standard_normal(3)
    ref = scipy.optimize.differential_evolution(f, bounds, updating='immediate', disp=False, maxiter=2000, polish=False)
    state_ref = np.random.get_state()
    np.random.seed(seed)
    np.random.standard_normal(3)
    got = ops.de_minimize(f, bounds, maxiter=2000, window=window)
    assert np.array_equal(got.x, ref.x) and got.fun == ref.fun
    assert (got.nit, got.nfev, got.converged) == (ref.nit, ref.nfev, bool(ref.success))
    assert _same_state(np.random.get_state(), state_ref)
    assert got.calls < got.nfev and got.calls * window >= got.nfev + got.discarded - window * (got.nit + 2)
    if f is not flat and window >= 16:
        assert got.calls * 3 < got.nfev                    # at least three retained evaluations per call


@pytest.mark.parametrize("maxiter", [2000, 3, 0])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_evolution_replays_scipy(case, maxiter):
    f, bounds, seed = CASES[case]
    np.random.seed(seed)
    np.random.standard_normal(3)               # a cached Gaussian and a position inside the state must survive the call
    ref = scipy.optimize.differential_evolution(f, bounds, updating='immediate', disp=False, maxiter=maxiter, polish=False)
    state_ref = np.random.get_state()
    np.random.seed(seed)
    np.random.standard_normal(3)
    got = ops.de_minimize(f, bounds, maxiter=maxiter)
    assert np.array_equal(got.x, ref.x) and got.fun == ref.fun
    assert (got.nit, got.nfev) == (ref.nit, ref.nfev)
    assert got.converged == bool(ref.success)
    assert got.population_size == len(ref.population)
    assert got.calls == got.nfev and got.discarded == 0
    assert _same_state(np.random.get_state(), state_ref)


@pytest.mark.parametrize("case", [0, 2, 3])
def test_polished_search_replays_scipy(case):
    """the reference's full call (polish=True is scipy's default): evolution in C++ + ops.de_polish == scipy, and the draws that
    follow (mu_star runs the search mustar_finding_trials times back to back) stay aligned"""
    f, bounds, seed = CASES[case]
    np.random.seed(seed)
    ref = [scipy.optimize.differential_evolution(f, bounds, updating='immediate', disp=False, maxiter=2000) for _ in range(2)]
    state_ref = np.random.get_state()
    np.random.seed(seed)
    got = [ops.de_polish(f, ops.de_minimize(f, bounds, maxiter=2000), bounds) for _ in range(2)]
    for r, g in zip(ref, got):
        assert np.array_equal(g.x, r.x) and g.fun == r.fun and g.nfev == r.nfev and g.nit == r.nit
    assert _same_state(np.random.get_state(), state_ref)


def test_other_settings_replay_scipy():
    f, bounds, seed = CASES[1]
    kw = dict(popsize=7, tol=1e-3, atol=1e-6, mutation=(0.3, 1.2), recombination=0.9)
    np.random.seed(seed)
    ref = scipy.optimize.differential_evolution(f, bounds, updating='immediate', disp=False, maxiter=60, polish=False, **kw)
    state_ref = np.random.get_state()
    np.random.seed(seed)
    got = ops.de_minimize(f, bounds, maxiter=60, **kw)
    assert np.array_equal(got.x, ref.x) and got.fun == ref.fun and (got.nit, got.nfev) == (ref.nit, ref.nfev)
    assert _same_state(np.random.get_state(), state_ref)


def test_objective_failures_surface():
    class Boom(Exception):
        pass

    def raises(x):
        raise Boom("objective failed")
    with pytest.raises(Boom):
        ops.de_minimize(raises, [(0, 1)] * 2, maxiter=5)
    calls = []

    def nan_later(x):
        calls.append(1)
        return float("nan") if len(calls) == 40 else sphere(x)
    with pytest.raises(PPBOError, match="NaN at evaluation 40"):
        ops.de_minimize(nan_later, [(0, 1)] * 2, maxiter=50)
    with pytest.raises(PPBOError):
        ops.de_minimize(sphere, [(0, np.inf)], maxiter=5)
    with pytest.raises(Boom):
        ops.de_minimize(raises, [(0, 1)] * 2, maxiter=5, window=8)
    with pytest.raises(PPBOError, match="window"):
        ops.de_minimize(sphere, [(0, 1)] * 2, maxiter=5, window=65)
    calls.clear()
    with pytest.raises(PPBOError, match="NaN"):
        ops.de_minimize(nan_later, [(0, 1)] * 2, maxiter=50, window=8)


def _random_objective(kind, D, rs):
    c, w, q = rs.rand(D), rs.rand(D) + 0.1, int(rs.randint(1, 4))
    if kind == 0:
        return lambda x: float(np.sum(w * (x - c) ** 2))
    if kind == 1:
        return lambda x: float(-np.exp(-np.sum((x - c) ** 2) / 0.05) - 0.3 * np.exp(-np.sum((x - 1 + c) ** 2) / 0.2))
    if kind == 2:
        return lambda x: float(np.round(np.sum(w * np.abs(x - c)), q))           # plateaus: ties between energies
    if kind == 3:
        return lambda x: float(np.sum(np.sin(7 * x + c)) + 0.1 * np.sum(x))
    return lambda x: float(np.floor(4 * np.sum(np.abs(x - c))))                  # few distinct values: ties everywhere


def test_random_problems_replay_scipy():
    """random dimensions, bounds, budgets, population sizes, windows and objectives (two of them full of ties, where '<=' against the
    member and against the best decide differently from '<'); a four-minute run of the same loop covered 3291 cases without a
    difference"""
    rs = np.random.RandomState(2024)
    for _ in range(80):
        D, kind = int(rs.randint(1, 8)), int(rs.randint(0, 5))
        f = _random_objective(kind, D, rs)
        bounds = [(0, 1)] * D if rs.rand() < 0.7 else [(float(-rs.rand() * 3), float(rs.rand() * 3 + 0.1)) for _ in range(D)]
        seed, maxiter = int(rs.randint(0, 2 ** 31 - 1)), int(rs.choice([2, 10, 40, 300]))
        popsize, window = int(rs.choice([3, 15])), int(rs.choice([1, 2, 3, 8, 32, 64]))
        np.random.seed(seed)
        ref = scipy.optimize.differential_evolution(f, bounds, updating='immediate', disp=False, maxiter=maxiter, polish=False,
                                                    popsize=popsize)
        state_ref = np.random.get_state()
        np.random.seed(seed)
        got = ops.de_minimize(f, bounds, maxiter=maxiter, popsize=popsize, window=window)
        assert np.array_equal(got.x, ref.x) and got.fun == ref.fun, (D, kind, bounds, seed, maxiter, popsize, window)
        assert (got.nit, got.nfev) == (ref.nit, ref.nfev)
        assert _same_state(np.random.get_state(), state_ref)
