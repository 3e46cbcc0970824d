"""Host-array convenience layer over the dense FP64 routines of the C ABI: numpy in, numpy out, arithmetic on the GPU.
Backs the reference-named helpers in src/misc.py (regularize_covariance, pd_inverse, inverse, is_positive_definite)."""
import numpy as np

from . import ops


def shrink_covariance(X, shrinkage):
    K = ops.to_dev(X).clone()
    return ops.shrink_inplace(K, shrinkage).cpu().numpy()


def _factor(M):
    A = ops.to_dev(M).clone()
    info, ws = ops.potrf_lower(A)
    return A, ws, info


def spd_inverse(M):
    if M.ndim != 2 or M.shape[0] != M.shape[1]:
        raise ValueError("expected a square matrix")
    A, ws, info = _factor(M)
    if info != 0:
        raise np.linalg.LinAlgError("%d-th leading minor of the array is not positive definite" % info)
    return ops.potri_lower(A, ws).cpu().numpy()


def spd_inverse_dev(A_dev):
    """device in / device out; A_dev is overwritten by its factor"""
    info, ws = ops.potrf_lower(A_dev)
    if info != 0:
        raise np.linalg.LinAlgError("%d-th leading minor of the array is not positive definite" % info)
    return ops.potri_lower(A_dev, ws)


def is_spd(M):
    return _factor(M)[2] == 0


def general_inverse(M):
    """A^-1 = (A^T A)^-1 A^T with the SPD route (conditioning is squared: a utility, not used on the hot path)."""
    A = ops.to_dev(M)
    At = A.t().contiguous()
    AtA = ops.gemm_nt(At, At)
    inv = spd_inverse_dev(AtA)
    return ops.gemm_nt(inv, A).cpu().numpy()          # (AtA)^-1 . A^T  == gemm_nt(inv, A): rows of A are K-contiguous
