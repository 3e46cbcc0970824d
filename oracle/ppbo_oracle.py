"""TEST INFRASTRUCTURE ONLY -- CPU (numpy/scipy, FP64) restatement of the PPBO hot path.

This file is the *checker* for the CUDA path in ``ppbo_b200/``.  It is never imported by
the product (``ppbo_b200/``, ``src/``); only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

It restates, function by function, what the reference (AaltoPML/PPBO, /root/reference)
computes on the path  GPModel.update_model -> next_query -> Hsampler.  Python-level loops
of the reference are vectorised (which makes this port *faster* than the real reference,
so speed-ups quoted against it are conservative), but the numerical recipe is kept:
expansion-form distances, SVD round trip + shrinkage, explicit SPD inverses, 200-point
Gauss-Hermite quadrature for the likelihood, scipy's ``trust-exact`` driver, numpy's
SVD-factor ``multivariate_normal`` fed from the legacy global RNG in reference call order.

PARITY PIN: the reference ships no tests or golden vectors, so this oracle is
pinned against the *executed* reference: ``oracle/make_golden.py`` runs the unmodified
reference (through ``oracle/ref_shim.py``) in the build container and stores its inputs and
outputs in ``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py`` checks every function
here against those files.
"""
import numpy as np
import scipy.linalg
import scipy.optimize
from scipy.special import ndtr

SQRT2 = np.sqrt(2.0)
SHRINKAGE = 1e-6            # GPModel.COVARIANCE_SHRINKAGE, src/gp_model.py:26
GH_POINTS = 200             # PPBO_settings.n_gausshermite_sample_points, src/ppbo_settings.py:52
EI_GRID_POINTS = 70         # hard-coded in src/acquisition.py:73,171


# --------------------------------------------------------------------------- kernels
def sqdist(X1, X2):
    """src/kernels.py:3-11 -- |x|^2 + |y|^2 - 2 x.y, clipped at 0."""
    n1 = np.einsum("ij,ij->i", X1, X1)
    n2 = np.einsum("ij,ij->i", X2, X2)
    return np.maximum(n1[:, None] + n2[None, :] - 2.0 * (X1 @ X2.T), 0.0)


def se_kernel(X1, X2, theta):
    """src/kernels.py:19-25 -- isotropic squared exponential, theta=(sigma, l, sigma_f)."""
    return theta[2] ** 2 * np.exp(-0.5 * sqdist(X1, X2) / theta[1] ** 2)


def rq_kernel(X1, X2, theta, alpha=2):
    """src/kernels.py:27-34 -- rational quadratic with alpha=2."""
    return theta[2] ** 2 * (1.0 + sqdist(X1, X2) / (2 * alpha * theta[1] ** 2)) ** (-alpha)


def camphor_copper_kernel(X1, X2, theta):
    """src/kernels.py:36-53 -- periodic (period 1) on dims 0,1,3,4,5; SE with l+0.05 on dim 2."""
    l, sf = theta[1], theta[2]
    K = np.full((X1.shape[0], X2.shape[0]), sf ** 2)
    for d in range(6):
        r = np.abs(X1[:, d][:, None] - X2[:, d][None, :])
        if d == 2:
            K = K * np.exp(-0.5 * r ** 2 / (l + 0.05) ** 2)
        else:
            K = K * np.exp(-2.0 * np.sin(np.pi * r) ** 2 / l ** 2)
    return K


KERNELS = {"SE_kernel": se_kernel, "RQ_kernel": rq_kernel, "camphor_copper_kernel": camphor_copper_kernel}


# --------------------------------------------------------------------------- linear algebra helpers
def regularize_covariance(K, shrinkage=SHRINKAGE, jitter=1e-7, svd_roundtrip=True):
    """src/misc.py:71-88 -- negative diagonal -> jitter; SVD round trip; sklearn shrunk_covariance
    == (1-s) K + s tr(K)/n I.  ``svd_roundtrip=False`` gives the closed form the CUDA path fuses."""
    K = np.array(K, dtype=float, copy=True)
    dg = np.diag(K).copy()
    dg[dg < 0] = jitter
    np.fill_diagonal(K, dg)
    if svd_roundtrip:
        u, s, vh = np.linalg.svd(K, full_matrices=False)
        K = np.matmul(np.matmul(u, np.diag(s)), vh)
    n = K.shape[0]
    mu = np.trace(K) / n
    K = (1.0 - shrinkage) * K
    K.flat[:: n + 1] += shrinkage * mu
    return K


def pd_inverse(M):
    """src/misc.py:96-100 -- solve(M, I) with the SPD (dposv) driver."""
    return scipy.linalg.solve(M, np.eye(M.shape[0]), assume_a="pos", overwrite_b=True)


def var2_normal_pdf(x):
    """src/misc.py:134-135 -- density of N(0, 2)."""
    return np.exp(-0.25 * np.square(x)) / np.sqrt(4.0 * np.pi)


_GH_CACHE = {}


def _gauss_hermite(n):
    if n not in _GH_CACHE:
        _GH_CACHE[n] = np.polynomial.hermite.hermgauss(n)
    return _GH_CACHE[n]


# --------------------------------------------------------------------------- likelihood terms
def deltas(f, Q, m, sigma):
    """Delta_qj = (f[pseudo j of set q] - f[winner of set q]) / sigma, shape (Q, m).
    Row layout: set q occupies rows q(m+1) .. q(m+1)+m, winner first (src/feedback_processing.py:121-123)."""
    F = np.asarray(f, dtype=float).reshape(Q, m + 1)
    return (F[:, 1:] - F[:, :1]) / sigma


def sum_phi(f, Q, m, sigma, order, quadrature=True):
    """src/gp_model.py:176-218 -- per comparison set: order 0: sum_j Phi~(Delta_j) (Gauss-Hermite 200,
    == Phi(Delta/sqrt2)); order 1: sum_j phi~(Delta_j); order 2: sum_j -Delta_j phi~(Delta_j)/2."""
    Dl = deltas(f, Q, m, sigma)
    if order == 0:
        if quadrature:
            t, w = _gauss_hermite(GH_POINTS)
            vals = ndtr(Dl[..., None] - SQRT2 * t) @ w / np.sqrt(np.pi)
        else:
            vals = ndtr(Dl / SQRT2)
        return vals.sum(axis=1)
    if order == 1:
        return var2_normal_pdf(Dl).sum(axis=1)
    if order == 2:
        return (-0.5 * Dl * var2_normal_pdf(Dl)).sum(axis=1)
    raise ValueError(order)


def T_value(f, Sigma_inv, Q, m, sigma, quadrature=True):
    """src/gp_model.py:221-226."""
    f = np.asarray(f, dtype=float).ravel()
    return -0.5 * f @ Sigma_inv @ f - sum_phi(f, Q, m, sigma, 0, quadrature).sum() / m


def lik_beta(f, Q, m, sigma):
    """Likelihood part of the gradient, src/gp_model.py:234-238."""
    ph = var2_normal_pdf(deltas(f, Q, m, sigma)) / (sigma * m)
    beta = np.empty((Q, m + 1))
    beta[:, 0] = ph.sum(axis=1)
    beta[:, 1:] = -ph
    return beta.ravel()


def T_grad(f, Sigma_inv, Q, m, sigma):
    """src/gp_model.py:228-240."""
    f = np.asarray(f, dtype=float).ravel()
    return -Sigma_inv @ f + lik_beta(f, Q, m, sigma)


def arrow_coeffs(f, Q, m, sigma):
    """a_qj = -Delta phi~(Delta) / (2 m sigma^2): W = -Lambda = B diag(a) B^T."""
    Dl = deltas(f, Q, m, sigma)
    return -0.5 * Dl * var2_normal_pdf(Dl) / (m * sigma ** 2)


def create_Lambda(f, Q, m, sigma):
    """src/gp_model.py:249-274 -- dense N x N star/arrow blocks."""
    a = arrow_coeffs(f, Q, m, sigma)
    N = Q * (m + 1)
    Lam = np.zeros((N, N))
    for q in range(Q):
        i = q * (m + 1)
        js = np.arange(i + 1, i + m + 1)
        Lam[js, js] = -a[q]
        Lam[i, i] = -a[q].sum()
        Lam[i, js] = a[q]
        Lam[js, i] = a[q]
    return Lam


def T_hessian(f, Sigma_inv, Q, m, sigma):
    """src/gp_model.py:242-247."""
    return -Sigma_inv + create_Lambda(f, Q, m, sigma)


# --------------------------------------------------------------------------- Laplace / MAP
def fmap_trust_exact(Sigma_inv, Q, m, sigma, f_initial, gtol=None):
    """src/gp_model.py:382-384 -- scipy trust-exact on -T with the dense Hessian (reference default gtol 1e-4)."""
    opts = {} if gtol is None else {"gtol": gtol}
    res = scipy.optimize.minimize(lambda f: -T_value(f, Sigma_inv, Q, m, sigma),
                                  np.asarray(f_initial, dtype=float).ravel(), method="trust-exact",
                                  jac=lambda f: -T_grad(f, Sigma_inv, Q, m, sigma),
                                  hess=lambda f: -T_hessian(f, Sigma_inv, Q, m, sigma), options=opts)
    return res.x, res


def fmap_tight(Sigma, Q, m, sigma, f_start, iters=60, tol=1e-13):
    """Tight stationary point of T near ``f_start``: Newton on the fixed point f = Sigma beta(f),
    never forming Sigma^{-1} (cond(Sigma) ~ 1e7+ makes the reference's own stopping point only
    ~1e-6 accurate).  Used as the 1e-6 yard-stick for the Laplace mode."""
    f = np.asarray(f_start, dtype=float).ravel().copy()
    N = f.size
    I = np.eye(N)
    for _ in range(iters):
        W = -create_Lambda(f, Q, m, sigma)
        b = W @ f + lik_beta(f, Q, m, sigma)
        f_new = Sigma @ np.linalg.solve(I + W @ Sigma, b)
        step = np.abs(f_new - f).max()
        f = f_new
        if step <= tol * max(1.0, np.abs(f).max()):
            break
    return f


def incidence(Q, m):
    """B [N x M]: column (q, j) has +1 at pseudo-observation j of set q and -1 at the set's winner, so that W = B diag(a) B^T
    (create_Lambda above) and B^T f are the differences the likelihood sees."""
    N, M = Q * (m + 1), Q * m
    B = np.zeros((N, M))
    for q in range(Q):
        for j in range(m):
            B[q * (m + 1) + 1 + j, q * m + j] = 1.0
            B[q * (m + 1), q * m + j] = -1.0
    return B


def chord_step_alpha(Sigma, f, a0, Q, m, sigma):
    """One chord step of the Newton iteration for Sigma^-1 f = beta(f) with FIXED clamped coefficients a0 >= 0 (those of the
    factor), in the (alpha, f) form of csrc/laplace.cu:  (I + W0 Sigma) alpha_new = W0 f + beta(f),  W0 = B a0 B^T,  solved through
    the M x M system  alpha_new = b - B s (I + s G s)^-1 s B^T Sigma b  (G = B^T Sigma B, s = sqrt(a0)).  Returns (alpha_new, f_new).
    Test infrastructure: restates the device iteration; the reference itself uses scipy trust-exact (src/gp_model.py:382-384)."""
    B = incidence(Q, m)
    s = np.sqrt(a0)
    G = B.T @ Sigma @ B
    b = B @ (a0 * (B.T @ f)) + lik_beta(f, Q, m, sigma)
    y = np.linalg.solve(np.eye(Q * m) + s[:, None] * G * s[None, :], s * (B.T @ (Sigma @ b)))
    alpha_new = b - B @ (s * y)
    return alpha_new, Sigma @ alpha_new


def chord_step_diff(G, gamma, d, a0, Q, m, sigma, threshold=0.0):
    """The same step carried by the M-vectors (gamma, d = G gamma) -- alpha = B gamma, d = B^T f:
        c = a0 d + g(d),   (I + s G s) y = s G c,   gamma_new = c - s y,   d_new = y / s  where a0 > threshold, else (G gamma_new)
    (from s G s y = s G c - y).  One product with G instead of two with Sigma.  Returns (gamma_new, d_new)."""
    s = np.sqrt(a0)
    g = -var2_normal_pdf(d / sigma) / (sigma * m)                  # beta = B g
    c = a0 * d + g
    y = np.linalg.solve(np.eye(Q * m) + s[:, None] * G * s[None, :], s * (G @ c))
    gamma_new = c - s * y
    big = a0 > threshold
    d_new = np.where(big, y / np.where(big, s, 1.0), G @ gamma_new)
    return gamma_new, d_new


def posterior_covariance(Sigma_inv, fMAP, Q, m, sigma):
    """src/gp_model.py:111-117."""
    Lam = create_Lambda(fMAP, Q, m, sigma)
    Pinv = Sigma_inv - Lam
    return Lam, Pinv, pd_inverse(Pinv)


# --------------------------------------------------------------------------- prediction
def mu_Sigma_pred(X, Xp, theta, kernel, Sigma_inv, fMAP, post_cov, svd_roundtrip=True):
    """src/gp_model.py:441-452 (same operation order as the reference)."""
    k = kernel(X, Xp, theta)
    mu = k.T.dot(Sigma_inv).dot(fMAP)
    Kss = regularize_covariance(kernel(Xp, Xp, theta), SHRINKAGE, svd_roundtrip=svd_roundtrip)
    A = Sigma_inv - Sigma_inv.dot(post_cov).dot(Sigma_inv)
    return mu, Kss - k.T.dot(A).dot(k)


def mu_pred(X, x, theta, kernel, Sigma_inv, fMAP):
    """src/gp_model.py:454-458."""
    k = kernel(X, np.asarray(x, dtype=float).reshape(1, -1), theta)
    return float((k.T @ Sigma_inv @ fMAP)[0])


# --------------------------------------------------------------------------- acquisition
def equispaced_alpha(alpha_min, alpha_max, P, noise=0.01):
    """src/feedback_processing.py:66-74 -- jittered linspace (consumes the global legacy RNG)."""
    eps_b = (alpha_max - alpha_min) * (noise / 2)
    eps_n = abs(alpha_max - alpha_min) * noise
    while True:
        a = np.linspace(alpha_min + eps_b, alpha_max - eps_b, num=P) + np.random.normal(0, eps_n, P)
        a = np.unique(np.clip(a, alpha_min, alpha_max))
        if len(a) == P:
            return a


def xi_grid_scaled(xi, x, P):
    """src/feedback_processing.py:47-108 with is_scaled=True, 'equispaced'."""
    a = equispaced_alpha(0.0, 1.0, P)
    return a[:, None] * np.asarray(xi, dtype=float)[None, :] + np.asarray(x, dtype=float)[None, :]


def mvn_svd_factor(cov):
    """numpy legacy multivariate_normal factor: x = mean + z @ (sqrt(s)[:,None] * vh)."""
    _, s, vh = np.linalg.svd(cov)
    return np.sqrt(s)[:, None] * vh


def sample_max(mu, cov, S):
    """f_max for S draws exactly as src/acquisition.py:78-80 / :175-177 (one draw per call)."""
    out = np.empty(S)
    for s in range(S):
        out[s] = np.max(np.random.multivariate_normal(mu, cov))
    return out


def EI_from_fmax(fmax, mustar):
    """src/acquisition.py:80-81."""
    return float(np.mean(np.maximum(fmax - mustar, 0.0)))


def varmax_from_fmax(fmax):
    """src/acquisition.py:178."""
    return float(np.mean(np.power(fmax - np.mean(fmax), 2)))


# --------------------------------------------------------------------------- random Fourier features
def rff_features(W, b, X, sigma_f):
    """src/random_fourier_sampler.py:45-47 -- phi(X) = sqrt(2 sf^2/F) cos(W X^T + b), shape F x n."""
    F = W.shape[0]
    return np.sqrt(2.0 * sigma_f ** 2 / F) * np.cos(W @ np.atleast_2d(X).T + np.asarray(b).reshape(F, 1))


def rff_jacobian(W, b, x, sigma_f):
    """src/random_fourier_sampler.py:51-53 -- d phi / d x, shape F x D."""
    F = W.shape[0]
    return -np.sqrt(2.0 * sigma_f ** 2 / F) * np.sin(W @ np.asarray(x, dtype=float) + np.asarray(b).reshape(F))[:, None] * W


def rff_S(omega, PhiX, Q, m, sigma, quadrature=True):
    """src/random_fourier_sampler.py:106-110."""
    f = PhiX.T @ omega
    return -0.5 * omega @ omega - sum_phi(f, Q, m, sigma, 0, quadrature).sum() / m


def _rff_diffs(PhiX, Q, m):
    F = PhiX.shape[0]
    P3 = PhiX.reshape(F, Q, m + 1)
    return P3[:, :, 1:] - P3[:, :, :1]          # F x Q x m : phi_j - phi_i


def rff_S_grad(omega, PhiX, Q, m, sigma):
    """src/random_fourier_sampler.py:112-116."""
    f = PhiX.T @ omega
    ph = var2_normal_pdf(deltas(f, Q, m, sigma)) / (sigma * m)
    return -omega - np.einsum("fqj,qj->f", _rff_diffs(PhiX, Q, m), ph)


def rff_S_hess_diag(omega, PhiX, Q, m, sigma):
    """Diagonal of src/random_fourier_sampler.py:118-122 (the reference Hessian *is* diagonal)."""
    f = PhiX.T @ omega
    Dl = deltas(f, Q, m, sigma)
    c = -0.5 * Dl * var2_normal_pdf(Dl) / (m * sigma ** 2)
    return -1.0 - np.einsum("fqj,qj->f", _rff_diffs(PhiX, Q, m) ** 2, c)


def rff_omega_map(PhiX, Q, m, sigma, omega0, gtol=None):
    """src/random_fourier_sampler.py:124-132."""
    opts = {} if gtol is None else {"gtol": gtol}
    res = scipy.optimize.minimize(lambda w: -rff_S(w, PhiX, Q, m, sigma), np.asarray(omega0, dtype=float),
                                  method="trust-exact", jac=lambda w: -rff_S_grad(w, PhiX, Q, m, sigma),
                                  hess=lambda w: -np.diag(rff_S_hess_diag(w, PhiX, Q, m, sigma)), options=opts)
    return res.x, res


def rff_eval_argmax(Omega, Phi_grid):
    """Batched form of the objective of src/random_fourier_sampler.py:166,170:
    Fs[S x P] = Omega[S x F] . Phi_grid[F x P]; per-sample max and first arg-max over P."""
    Fs = Omega @ Phi_grid
    idx = np.argmax(Fs, axis=1)
    return Fs[np.arange(Fs.shape[0]), idx], idx.astype(np.int32)


# --------------------------------------------------------------------------- device RNG (not in the reference)
def philox4x32_10(counter, key):
    """Philox4x32-10 of Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3" (SC'11), the
    published algorithm (Random123 v1.x constants).  counter: (n, 4) uint32, key: (2,) uint32 -> (n, 4) uint32.
    The reference draws from numpy's global MT19937; this generator only backs the CUDA path's own draws when the
    caller does not inject normals, so its oracle is the published algorithm (known-answer vectors in the tests)."""
    c = np.array(counter, dtype=np.uint64).reshape(-1, 4)
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def philox_normals(seed, stream_id, offset, n):
    """Standard normals offset .. offset+n-1 of the stream: counter i = (i_lo, i_hi, stream_id, 0) yields numbers 2i, 2i+1 by
    Box-Muller on two 53-bit uniforms u = ((a >> 5) 2^26 + (b >> 6) + 1/2) 2^-53."""
    first, last = offset // 2, (offset + n - 1) // 2
    idx = np.arange(first, last + 1, dtype=np.uint64)
    ctr = np.stack([idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.full_like(idx, stream_id), np.zeros_like(idx)], axis=1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.float64)
    u1 = (np.floor(r[:, 0] / 32) * 67108864.0 + np.floor(r[:, 1] / 64) + 0.5) / 9007199254740992.0
    u2 = (np.floor(r[:, 2] / 32) * 67108864.0 + np.floor(r[:, 3] / 64) + 0.5) / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.stack([rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)], axis=1).ravel()
    return z[offset - 2 * first: offset - 2 * first + n]


def rff_sample_omega(omega_map, hess_diag, Z):
    """src/random_fourier_sampler.py:134-137,207-213 with the (diagonal) Laplace covariance 1 / -hess_diag and injected
    standard normals Z [S x F]."""
    return np.asarray(omega_map)[None, :] + np.asarray(Z) / np.sqrt(-np.asarray(hess_diag))[None, :]


# --------------------------------------------------------------------------- INT8 error-free splitting (csrc/ozaki.cu)
# Restatement of the arithmetic of the tcgen05 path of the sampling contraction (batched objective of
# Hsampler.return_xstar, src/random_fourier_sampler.py:166,170).  Integer arithmetic is exact, so the CUDA kernels must
# reproduce these functions bit for bit.
OZAKI_KB = 64        # bytes of K per digit-plane tile


def ozaki_rowscale(X):
    """scale[r] = 2^(E+2), max_k |X[r,k]| = m 2^E with m in [1/2, 1); 1 for an all-zero row."""
    amax = np.abs(np.asarray(X, dtype=np.float64)).max(axis=1)
    _, e = np.frexp(amax)
    return np.where(amax > 0, np.ldexp(1.0, e + 2), 1.0)


def ozaki_digits(X, slices):
    """(digits [slices, rows, K] int8, scale [rows]):  X[r,k] ~= scale[r] * sum_i digits[i,r,k] 256^-(i+1).
    One rounding (half-to-even) to a fixed-point integer of 8*slices bits, then balanced base-256 digits in [-128, 127]."""
    X = np.asarray(X, dtype=np.float64)
    scale = ozaki_rowscale(X)
    Y = np.rint(X * np.ldexp(1.0 / scale, 8 * slices)[:, None]).astype(np.int64)
    digits = np.empty((slices,) + X.shape, dtype=np.int8)
    for s in range(slices - 1, -1, -1):
        d = ((Y + 128) & 255) - 128
        Y = (Y - d) >> 8
        digits[s] = d.astype(np.int8)
    assert not Y.any()
    return digits, scale


def ozaki_planes(digits, tile_rows):
    """byte image of the digit planes as the tensor-core operand reads them: [row tile][k block][slice][tile_rows x 64 B] with
    the 8 x 16 B core matrices of the no-swizzle K-major UMMA layout (8-row groups 512 B apart, K chunks 128 B apart)."""
    KS, rows, K = digits.shape
    NT, KBLK = -(-rows // tile_rows), -(-K // OZAKI_KB)
    pad = np.zeros((KS, NT * tile_rows, KBLK * OZAKI_KB), dtype=np.int8)
    pad[:, :rows, :K] = digits
    # (s, rt, rg, r8, kb, c, byte) -> (rt, kb, s, rg, c, r8, byte)
    t = pad.reshape(KS, NT, tile_rows // 8, 8, KBLK, OZAKI_KB // 16, 16).transpose(1, 4, 0, 2, 5, 3, 6)
    return np.ascontiguousarray(t).reshape(-1)


def ozaki_matmul(A, B, slices):
    """C = A . B^T through `slices` digit planes per operand: exact integer accumulators acc_d = sum_{i+j=d} A_i . B_j^T for
    d < slices, recombined as (hi + lo) in INT64 -> FP64 with ONE rounding, then the two power-of-two row scales."""
    da, sa = ozaki_digits(A, slices)
    db, sb = ozaki_digits(B, slices)
    g1 = min(3, slices)
    hi = np.zeros((A.shape[0], B.shape[0]), dtype=np.int64)
    lo = np.zeros_like(hi)
    fa, fb = da.astype(np.float64), db.astype(np.float64)          # float64 BLAS on small integers is exact (< 2^53)
    for d in range(slices):
        acc = np.zeros(hi.shape)
        for i in range(d + 1):
            acc += fa[i] @ fb[d - i].T
        acc = acc.astype(np.int64)
        assert np.abs(acc).max() < 2 ** 31
        if d < g1:
            hi = hi * 256 + acc
        else:
            lo = lo * 256 + acc
    v = lo.astype(np.float64) * np.ldexp(1.0, -8 * (slices + 1)) + hi.astype(np.float64) * np.ldexp(1.0, -8 * (g1 + 1))
    return v * sb[None, :] * sa[:, None]


def ozaki_eval_argmax(Omega, PhiT, slices):
    """per-sample max / first arg-max of ozaki_matmul(Omega, PhiT): what ppbo_ozaki_rowmax returns for one grid"""
    Fs = ozaki_matmul(Omega, PhiT, slices)
    return Fs.max(axis=1), Fs.argmax(axis=1).astype(np.int32), Fs
