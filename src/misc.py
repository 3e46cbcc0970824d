"""Linear-algebra and density helpers with the reference's names (src/misc.py).  Array helpers that the hot path calls on
N x N matrices (regularize_covariance, pd_inverse, inverse) run on the GPU through the C ABI; scalar/index helpers
(alpha_bounds, hypercube_corners, pdf of a scalar) are host glue exactly as in the reference."""
import itertools
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ppbo_b200 import device_linalg as _dl  # noqa: E402


def alpha_bounds(xi, lower, upper):
    """Range of alpha for which alpha * xi stays inside the box [lower, upper] (src/misc.py:27-61).  Host glue (D numbers)."""
    xi, lower, upper = (np.asarray(v, dtype=float) for v in (xi, lower, upper))
    pos, neg = xi > 0, xi < 0
    lows = np.concatenate([lower[pos] / xi[pos], upper[neg] / xi[neg]])
    highs = np.concatenate([lower[neg] / xi[neg], upper[pos] / xi[pos]])
    alpha_lower = lows.max() if lows.size else -np.inf
    alpha_upper = highs.min() if highs.size else np.inf
    if alpha_lower > alpha_upper:
        print("Error: alpha_min > alpha_max!")
    if alpha_lower == -np.inf:
        print("Error: alpha_min is -infinity!")
    if alpha_upper == np.inf:
        print("Error: alpha_max is infinity!")
    return alpha_lower, alpha_upper


def regularize_covariance(X, reg_level=1e-4, pos_diag=True, jitter=1e-7):
    """(1 - s) X + s tr(X)/n I after forcing a positive diagonal (src/misc.py:71-88).  The reference's SVD round trip
    u diag(s) vh is the identity map up to round-off and is not reproduced."""
    X = np.array(X, dtype=np.float64, copy=True)
    if pos_diag:
        dg = np.diag(X).copy()
        dg[dg < 0] = jitter
        np.fill_diagonal(X, dg)
    return _dl.shrink_covariance(X, reg_level)


def inverse(matrix):
    """general inverse (src/misc.py:91-93); unused on the hot path -- the SPD route below is what the model calls."""
    return _dl.general_inverse(np.asarray(matrix, dtype=np.float64))


def pd_inverse(matrix):
    """inverse of a symmetric positive definite matrix (src/misc.py:96-100, LAPACK dposv in the reference): blocked Cholesky
    and two triangular sweeps on the GPU.  Raises numpy.linalg.LinAlgError when the matrix is not positive definite, as
    scipy.linalg.solve(assume_a='pos') does."""
    return _dl.spd_inverse(np.asarray(matrix, dtype=np.float64))


def det(matrix, regularization_level=0):
    """determinant through the Cholesky/LU log-determinant (src/misc.py:103-112); host, unused on the hot path."""
    M = np.asarray(matrix, dtype=float)
    M = M + np.eye(len(M)) * np.max(np.diag(M)) * regularization_level
    sign, logdet = np.linalg.slogdet(M)
    return sign * np.exp(logdet)


def pseudo_det(matrix):
    """product of (at most 300) eigenvalues with positive real part (src/misc.py:114-118); host, unused on the hot path."""
    ev = np.linalg.eig(np.asarray(matrix, dtype=float))[0]
    ev = ev[np.real(ev) > 1e-12][:300]
    return np.abs(np.prod(ev))


def is_positive_definite(M):
    """True iff the GPU Cholesky of M succeeds (src/misc.py:120-126)."""
    ok = _dl.is_spd(np.asarray(M, dtype=np.float64))
    if not ok:
        print('Function is_positive_definite: Matrix is not positive definite!')
    return ok


def std_normal_pdf(x):
    return np.exp(-0.5 * np.square(x)) / np.sqrt(2 * np.pi)


def var2_normal_pdf(x):
    """density of N(0, 2) (src/misc.py:134-135)"""
    return np.exp(-0.25 * np.square(x)) / np.sqrt(4 * np.pi)


def hypercube_corners(bounds):
    return np.array(list(itertools.product(*[(b[0], b[1]) for b in bounds])))
