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


def general_inverse(M, refine=2):
    """A^-1 from the normal equations, X0 = (A^T A)^-1 A^T with the SPD route, followed by `refine` Newton-Schulz steps
    X <- X + X (I - A X).  The normal equations square the condition number (error ~ cond(A)^2 eps); each step squares the
    residual ||I - A X||, so two steps return the cond(A) eps accuracy of an LU-based inverse whenever cond(A)^2 eps < 1.
    A utility behind misc.inverse (src/misc.py:91-93), not used on the hot path."""
    import torch
    A = ops.to_dev(M)
    n = A.shape[0]
    At = A.t().contiguous()
    AtA = ops.gemm_nt(At, At)
    inv = spd_inverse_dev(AtA)
    X = ops.gemm_nt(inv, A)                            # (AtA)^-1 . A^T  == gemm_nt(inv, A): rows of A are K-contiguous
    for _ in range(int(refine)):
        R = torch.eye(n, dtype=A.dtype, device=A.device)
        ops.gemm_nt(A, X.t().contiguous(), C=R, alpha=-1.0, beta=1.0)       # R = I - A X
        if not bool(torch.isfinite(R).all()) or float(R.abs().sum(dim=1).max()) >= 1.0:
            break                                      # outside the convergence region: keep the unrefined inverse
        Xn = X.clone()                                 # the product reads X while it writes C: no aliasing
        X = ops.gemm_nt(X, R.t().contiguous(), C=Xn, alpha=1.0, beta=1.0)   # X + X R
    return X.cpu().numpy()
