"""INT8 error-free splitting of the sampling contraction (csrc/ozaki.cu, tcgen05.mma.kind::i8).
Host tests pin the oracle restatement (oracle/ppbo_oracle.py ozaki_*) against plain FP64; GPU tests require the CUDA path to
reproduce the oracle BIT FOR BIT (integer arithmetic is exact) and to agree with the FP64 DMMA kernel on max / arg-max."""
import numpy as np
import pytest

from oracle import ppbo_oracle as O

torch = pytest.importorskip("torch")


def _operands(S, F, P, seed=0, sigma_f=0.5):
    rng = np.random.RandomState(seed)
    Omega = 0.3 * rng.randn(S, F) + 0.05 * rng.randn(F)[None, :]
    PhiT = np.sqrt(2.0 * sigma_f ** 2 / F) * np.cos(3.0 * rng.randn(P, F) + rng.uniform(0, 2 * np.pi, F)[None, :])
    return Omega, PhiT


# ----------------------------------------------------------------------------------------------- host
@pytest.mark.parametrize("slices,tol", [(5, 2e-10), (6, 1e-12), (7, 1e-14)])
def test_oracle_ozaki_accuracy(slices, tol):
    Omega, PhiT = _operands(150, 1000, 96)
    ref = Omega @ PhiT.T
    C = O.ozaki_matmul(Omega, PhiT, slices)
    assert np.abs(C - ref).max() <= tol * np.abs(ref).max()
    # a-priori bound: dropped digit pairs (i + j >= slices) + one rounding of every operand entry to 8 slices - 2 bits
    sa, sb = O.ozaki_rowscale(Omega), O.ozaki_rowscale(PhiT)
    bound = (slices + 2) * Omega.shape[1] * 2.0 ** (-8 * slices - 2) * sa[:, None] * sb[None, :]
    assert np.all(np.abs(C - ref) <= bound + 1e-15 * np.abs(ref).max())


def test_oracle_ozaki_digits_roundtrip_and_range():
    rng = np.random.RandomState(1)
    X = rng.randn(40, 130) * np.exp(rng.randn(40, 1) * 5)
    X[3] = 0.0
    X[5, 7] = -X[5].__abs__().max() * 1.5          # a negative row maximum
    for ks in (2, 5, 6, 7):
        d, s = O.ozaki_digits(X, ks)
        assert d.dtype == np.int8 and np.all(np.log2(s) == np.round(np.log2(s)))
        rec = sum(d[i].astype(np.float64) * 256.0 ** -(i + 1) for i in range(ks)) * s[:, None]
        assert np.abs(rec - X).max() <= (s * 2.0 ** (-8 * ks - 1)).max()
        assert np.all(np.abs(d[0].astype(int)) <= 65)      # |x| / scale < 1/4 leaves headroom for the carry
    d, s = O.ozaki_digits(X, 7)                            # 54 bits: exact for entries within 2^-1 of the row maximum
    big = np.abs(X) >= 0.5 * np.abs(X).max(axis=1, keepdims=True)
    rec = sum(d[i].astype(np.float64) * 256.0 ** -(i + 1) for i in range(7)) * s[:, None]
    assert np.array_equal(rec[big], X[big])


def test_oracle_ozaki_exact_power_of_two_scaling_and_layout():
    Omega, PhiT = _operands(20, 100, 24, seed=3)
    C = O.ozaki_matmul(Omega, PhiT, 6)
    assert np.array_equal(O.ozaki_matmul(4.0 * Omega, PhiT, 6), 4.0 * C)          # linear in powers of two, exactly
    assert np.array_equal(O.ozaki_matmul(Omega, 0.125 * PhiT, 6), 0.125 * C)
    d, _ = O.ozaki_digits(Omega, 6)
    img = O.ozaki_planes(d, 128)
    assert img.size == 1 * 2 * 6 * 128 * 64
    # element (r=13, k=70) of slice 2: row tile 0, k block 1, row group 1, chunk 0, row 5 of the group, byte 6
    off = ((0 * 2 + 1) * 6 + 2) * 128 * 64 + (13 // 8) * 512 + ((70 % 64) // 16) * 128 + (13 % 8) * 16 + (70 % 16)
    assert img[off] == d[2, 13, 70]


# ----------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ppbo_b200 import ops as _ops
    return _ops


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("slices", [5, 6, 7])
def test_slice_kernel_bit_exact(ops, slices):
    rng = np.random.RandomState(2)
    X = rng.randn(300, 200) * np.exp(rng.randn(300, 1) * 3)
    X[17] = 0.0
    planes, scale = ops.ozaki_slice(ops.to_dev(X), 0, slices)
    d, s = O.ozaki_digits(X, slices)
    assert np.array_equal(_np(scale)[:300], s) and np.all(_np(scale)[300:] == 1.0)
    assert np.array_equal(_np(planes), O.ozaki_planes(d, 128))
    Xb = rng.randn(3, 100, 200)                                            # batched B operand, 64-row tiles
    planes, scale = ops.ozaki_slice(ops.to_dev(Xb), 1, slices)
    img = np.concatenate([O.ozaki_planes(O.ozaki_digits(Xb[b], slices)[0], 64) for b in range(3)])
    assert np.array_equal(_np(planes), img)
    assert np.array_equal(_np(scale).reshape(3, 128)[:, :100], np.stack([O.ozaki_rowscale(Xb[b]) for b in range(3)]))


@pytest.mark.gpu
@pytest.mark.parametrize("slices", [5, 6, 7])
@pytest.mark.parametrize("S,F,P,B", [(128, 64, 64, 1), (300, 200, 100, 2), (129, 1000, 70, 3), (1000, 333, 257, 2)])
def test_rowmax_i8_bit_exact_vs_oracle(ops, slices, S, F, P, B):
    Omega, _ = _operands(S, F, 8, seed=S)
    grids = np.stack([_operands(8, F, P, seed=10 * b + P)[1] for b in range(B)])
    Omega[S // 2] = 0.0
    fmax, arg, full = ops.rff_eval_argmax_i8(ops.to_dev(Omega), ops.to_dev(grids), slices=slices, want_full=True)
    torch.cuda.synchronize()
    for b in range(B):
        mx, am, Fs = O.ozaki_eval_argmax(Omega, grids[b], slices)
        assert np.array_equal(_np(full)[b], Fs)
        assert np.array_equal(_np(fmax)[b], mx)
        assert np.array_equal(_np(arg)[b], am)


@pytest.mark.gpu
def test_rowmax_i8_first_argmax_on_ties_and_launch_shapes(ops):
    """Equal maxima in the two column halves of a tile (two epilogue warps per row), in different tiles and in different column
    groups resolve to the LOWEST column, as np.argmax does; other launch shapes (tuning key 14: SMs left free for the overlapped
    one-GPU pipeline, or one CTA per work item) return the same bits."""
    from ppbo_b200 import _lib
    S, F, P, B = 300, 128, 200, 3
    Omega, _ = _operands(S, F, 8, seed=3)
    g = _operands(8, F, P, seed=4)[1]
    grids = np.stack([g, g, g]).copy()
    big = 3.0 * g[5]
    grids[0][5] = grids[0][40] = grids[0][70] = big          # half 0 / half 1 of tile 0, half 0 of tile 1   -> 5
    grids[1][40] = grids[1][70] = big                         # half 1 of tile 0 against half 0 of tile 1      -> 40
    grids[2][33] = grids[2][199] = big                        # half 1 of tile 0 against tile 3                -> 33
    lib = _lib.load()
    outs = []
    for key, groups in ((0, 0), (-1, 0), (0, 2), (100, 4)):     # key 4: number of column-tile groups (partial maxima merged afterwards)
        lib.ppbo_set_tuning(14, key)
        lib.ppbo_set_tuning(4, groups)
        try:
            fmax, arg, full = ops.rff_eval_argmax_i8(ops.to_dev(Omega), ops.to_dev(grids), slices=6, want_full=True)
            torch.cuda.synchronize()
        finally:
            lib.ppbo_set_tuning(14, 0)
            lib.ppbo_set_tuning(4, 0)
        outs.append((_np(fmax), _np(arg), _np(full)))
    for b, first in enumerate((5, 40, 33)):
        mx, am, Fs = O.ozaki_eval_argmax(Omega, grids[b], 6)
        assert (am == first).sum() > 20                  # the tie really decides the arg-max for many samples
        for fmax, arg, full in outs:
            assert np.array_equal(full[b], Fs) and np.array_equal(fmax[b], mx) and np.array_equal(arg[b], am)


@pytest.mark.gpu
def test_rowmax_i8_bench_shape_vs_fp64_oracle(ops):
    """the bench's contraction shape (F = 1000, P = 1024) against the ORACLE's FP64 product (numpy, host): values to the error bound
    of the digit count, arg-max identical wherever the oracle's gap to the runner-up exceeds that bound (tie rule), and the FP64
    DMMA kernel held to the same standard"""
    S, F, P, B = 4096, 1000, 1024, 3
    Omega, _ = _operands(S, F, 8, seed=5)
    grids = np.stack([_operands(8, F, P, seed=50 + b)[1] for b in range(B)])
    Od, Gd = ops.to_dev(Omega), ops.to_dev(grids)
    ref_max, ref_arg, ref_gap = [], [], []
    for b in range(B):
        Fs = Omega @ grids[b].T
        idx = Fs.argmax(axis=1)
        top = Fs[np.arange(S), idx]
        Fs[np.arange(S), idx] = -np.inf
        ref_max.append(top), ref_arg.append(idx), ref_gap.append(top - Fs.max(axis=1))
    ref_max, ref_arg, ref_gap = np.array(ref_max), np.array(ref_arg), np.array(ref_gap)
    scale = float(np.abs(ref_max).max())
    f64max, f64arg, _ = ops.rff_eval_argmax(Od, Gd)
    for name, fmax, arg, tol in [("f64", f64max, f64arg, 1e-12)] + [
            ("i8-%d" % k, *ops.rff_eval_argmax_i8(Od, Gd, slices=k)[:2], t) for k, t in ((5, 1e-9), (6, 1e-11), (7, 1e-12))]:
        fmax, arg = _np(fmax), _np(arg)
        assert np.abs(fmax - ref_max).max() <= tol * scale, name
        decided = ref_gap > 4 * tol * scale
        assert decided.mean() > 0.999, name
        assert np.array_equal(arg[decided], ref_arg[decided]), name     # identical arg-max; below the bound either index is one
        assert np.all(fmax[~decided] >= ref_max[~decided] - tol * scale), name
    # repeatability: bit-identical across launches
    a1 = ops.rff_eval_argmax_i8(Od, Gd, slices=6)[0]
    a2 = ops.rff_eval_argmax_i8(Od, Gd, slices=6)[0]
    assert torch.equal(a1, a2)


@pytest.mark.gpu
@pytest.mark.parametrize("S,F,sample0", [(128, 64, 0), (300, 1000, 0), (257, 333, 7), (1000, 96, 12345), (8, 65, 1)])
def test_fused_draw_and_slice_bit_identical(ops, S, F, sample0):
    """ppbo_ozaki_sample_slice (draws generated in shared memory, never written to HBM) against ppbo_rff_sample_omega followed by
    ppbo_ozaki_slice: identical digit planes and row scales, for ragged S, odd F and an offset into the Philox stream"""
    rng = np.random.RandomState(F)
    om = ops.to_dev(rng.randn(F) * 0.3)
    hd = ops.to_dev(-(1.0 + 5 * rng.rand(F)))
    for slices in (5, 6, 7):
        Omega = ops.rff_sample_omega(om, hd, S, seed=42, stream_id=3, sample0=sample0)
        p0, s0 = ops.ozaki_slice(Omega, 0, slices)
        p1, s1 = ops.ozaki_sample_slice(om, hd, S, seed=42, stream_id=3, sample0=sample0, slices=slices)
        assert torch.equal(s0, s1)
        assert torch.equal(p0, p1)
