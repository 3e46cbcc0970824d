// ppbo_b200 -- dense FP64 linear algebra on sm_100a: GEMM launchers, blocked Cholesky, triangular solves, GEMV.
// Replaces the LAPACK/BLAS calls the reference reaches through numpy/scipy (dpotrf/dposv under
// scipy.linalg.solve(assume_a='pos') in src/misc.py:96-100 and under scipy's trust-exact used by
// src/gp_model.py:382-384; dgemm/dgemv under np.dot in src/gp_model.py:441-458).
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "../../include/ppbo_b200.h"
#include "gemm_f64.cuh"
#include "linalg.cuh"

namespace ppbo {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------- GEMM launchers
using CfgBig = GemmCfg<128, 128, 2, 4, 4, 2>;      // 256 threads, warp tile 64x32, 160 KB smem, 1 CTA / SM
using CfgBigU = GemmCfg<128, 128, 2, 4, 4, 1>;     // same, 8-byte copies for odd leading dimensions
using CfgWide = GemmCfg<64, 128, 2, 4, 4, 2>;      // in-place row-panel updates: one CTA owns whole rows (BN >= K)
using CfgWideU = GemmCfg<64, 128, 2, 4, 4, 1>;
using CfgSmall = GemmCfg<64, 64, 2, 2, 4, 2>;      // 128 threads, warp tile 32x32, 80 KB smem, 2 CTAs / SM
using CfgSmallU = GemmCfg<64, 64, 2, 2, 4, 1>;
// experimental configurations reachable through ppbo_gemm_nt_cfg (scripts/ubench_ops.py)
using CfgTall3 = GemmCfg<128, 64, 4, 2, 3, 2, 2>;  // 256 threads, warp tile 32x32, 92 KB smem, 2 CTAs / SM
using CfgSq16 = GemmCfg<128, 128, 4, 4, 4, 2>;     // 512 threads, warp tile 32x32, 160 KB smem, 1 CTA / SM
using CfgSq3 = GemmCfg<128, 128, 2, 4, 3, 2>;      // CfgBig with 3 stages (120 KB)
using CfgTall4 = GemmCfg<128, 64, 2, 2, 4, 2>;     // 128 threads, warp tile 64x32, 123 KB, 1 CTA / SM

template <class Cfg>
static int set_smem_attr_store() {
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] {
        err = cudaFuncSetAttribute(gemm_nt_store_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    });
    PPBO_CUDA_CHECK(err);
    return PPBO_OK;
}
template <class Cfg>
static int set_smem_attr_rowmax() {
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] {
        err = cudaFuncSetAttribute(gemm_nt_rowmax_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    });
    PPBO_CUDA_CHECK(err);
    return PPBO_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool operands_vec2(const GemmOperands& g) {
    return aligned16(g.A) && aligned16(g.B) && g.lda % 2 == 0 && g.ldb % 2 == 0 && g.strideA % 2 == 0 && g.strideB % 2 == 0 &&
           g.strideA2 % 2 == 0 && g.strideB2 % 2 == 0;
}

PinnedReadback& readback() {
    static thread_local PinnedReadback r;
    return r;
}

cudaError_t malloc_async(void** p, size_t bytes, cudaStream_t st) {
    static std::once_flag once[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64)
        std::call_once(once[dev], [dev] {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            (void)cudaGetLastError();
        });
    return cudaMallocAsync(p, bytes, st);
}

static thread_local bool g_launch_pdl = false;      // set by potrf_lower around the launches of its critical path

template <class Cfg>
static int launch_store_cfg(const GemmOperands& g, StoreEpilogue ep, int batch, cudaStream_t st) {
    int rc = set_smem_attr_store<Cfg>();
    if (rc) return rc;
    const int tm = ceil_div(g.M, Cfg::BM), tn = ceil_div(g.N, Cfg::BN);
    dim3 grid;
    if (ep.lower_only) {
        constexpr int R = (Cfg::BM >= Cfg::BN) ? Cfg::BM / Cfg::BN : 1;
        grid = dim3((unsigned)((long long)R * tm * (tm + 1) / 2), 1, batch);
    } else {
        grid = dim3(tm, tn, batch);
    }
    if (g_launch_pdl) {
        // critical path of the blocked Cholesky: this kernel may become resident while its predecessor in the stream is still
        // running (it waits at griddepcontrol.wait before touching memory), which hides launch and scheduling latency
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ++g_launch_count;
        PPBO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_nt_store_kernel<Cfg>, g, ep));
        return PPBO_OK;
    }
    PPBO_CL gemm_nt_store_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(g, ep);
    PPBO_LAUNCH_CHECK();
    return PPBO_OK;
}

int launch_gemm_nt(const GemmOperands& g, const StoreEpilogue& ep_in, int batch, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return PPBO_OK;
    StoreEpilogue ep = ep_in;
    ep.vec_ok = aligned16(ep.C) && ep.ldc % 2 == 0 && ep.strideC % 2 == 0 && ep.strideC2 % 2 == 0;
    const bool v2 = operands_vec2(g);
    // small problems: more, smaller CTAs so the 148 SMs see work
    const long long big_tiles = (long long)ceil_div(g.M, 128) * ceil_div(g.N, 128) * batch;
    const bool small = !ep.lower_only && big_tiles < PPBO_SM_COUNT;
    if (ep.lower_only) PPBO_REQUIRE(g.M == g.N, "lower_only needs a square output");
    if (ep.in_place) {   // C aliases A: the CTA must read all of its rows (every k) before it stores -> BN >= N == K
        PPBO_REQUIRE(g.N <= 128 && g.K <= 128 && !ep.lower_only, "in-place update needs N, K <= 128");
        return v2 ? launch_store_cfg<CfgWide>(g, ep, batch, st) : launch_store_cfg<CfgWideU>(g, ep, batch, st);
    }
    if (small) return v2 ? launch_store_cfg<CfgSmall>(g, ep, batch, st) : launch_store_cfg<CfgSmallU>(g, ep, batch, st);
    if (v2) return launch_store_cfg<CfgTall3>(g, ep, batch, st);   // 2 CTAs / SM: prologue/epilogue overlap
    return v2 ? launch_store_cfg<CfgBig>(g, ep, batch, st) : launch_store_cfg<CfgBigU>(g, ep, batch, st);
}

template <class Cfg>
static int launch_rowmax_cfg(const GemmOperands& g, const RowMaxEpilogue& ep_in, int batch, cudaStream_t st) {
    int rc = set_smem_attr_rowmax<Cfg>();
    if (rc) return rc;
    RowMaxEpilogue ep = ep_in;
    const int tiles_m = ceil_div(g.M, Cfg::BM);
    // A rows of one group ~48 MB (L2 is 126 MB; the B operand of the 2-3 batch entries in flight takes another ~25 MB)
    const long long tile_bytes = (long long)Cfg::BM * max(g.K, 1) * 8;
    int group_m = (int)std::max<long long>(8, std::min<long long>(tiles_m, (48LL << 20) / tile_bytes));
    if (g.strideA != 0) group_m = tiles_m;      // A differs per batch entry (exact-GP draws): nothing to reuse across entries
    ep.batch = batch;
    ep.group_m = group_m;
    const int groups = ceil_div(tiles_m, group_m);
    dim3 grid((unsigned)((long long)groups * group_m * batch), 1, 1);
    PPBO_CL gemm_nt_rowmax_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(g, ep);
    PPBO_LAUNCH_CHECK();
    return PPBO_OK;
}

int g_tuning[16] = {0};     // [0]: row-max GEMM tile configuration (0 = default)

int launch_gemm_nt_rowmax(const GemmOperands& g, const RowMaxEpilogue& ep, int batch, cudaStream_t st) {
    if (g.M <= 0 || batch <= 0) return PPBO_OK;
    PPBO_REQUIRE(g.N > 0 && g.K >= 0, "empty grid");
    const bool v2 = operands_vec2(g);
    const bool small = (long long)ceil_div(g.M, 128) * batch < PPBO_SM_COUNT;
    // default for large problems: 128 x 64 tiles, 2 CTAs / SM (32.0 TFLOP/s on the Ackley-20D contraction vs 29.7 for 128 x 128)
    if (v2 && !small && g_tuning[0] == 1) return launch_rowmax_cfg<CfgBig>(g, ep, batch, st);
    if (v2 && !small && g_tuning[0] == 3) return launch_rowmax_cfg<CfgSq16>(g, ep, batch, st);
    if (v2 && !small) return launch_rowmax_cfg<CfgTall3>(g, ep, batch, st);
    if (small) return v2 ? launch_rowmax_cfg<CfgSmall>(g, ep, batch, st) : launch_rowmax_cfg<CfgSmallU>(g, ep, batch, st);
    return v2 ? launch_rowmax_cfg<CfgBig>(g, ep, batch, st) : launch_rowmax_cfg<CfgBigU>(g, ep, batch, st);
}

// ------------------------------------------------------------------------------------------- row panel x 128 x 128 block
// out[r][0:128] = alpha * sum_k A[r][k] Bm[j][k] + beta * C[r][j]   for r < M, with N = K = 128 fixed.
// The two GEMMs on the critical path of every Cholesky step have this shape (panel L21 = A21 inv(L11)^T, in place; look-ahead
// A22[:, 0:128] -= L21 L21[0:128]^T).  The general kernel spends ~15-20 us on each (3-stage pipeline fill, 8 barrier-gated
// k-tiles of 16, C read only in the epilogue) for 4 us of DMMA work per SM.  Here the whole 128 x 128 block and the CTA's R rows go
// to shared memory in four 32-column cp.async groups (compute on group g while g+1.. are in flight), the C fragments are
// requested before the main loop, and rows are padded to 132 doubles (== 4 mod 16: the 8-byte fragment loads of a half-warp
// hit 16 distinct bank pairs).  `out` may alias A (each CTA has read its own rows completely before it stores).
constexpr int RP_LD = 132, RP_THREADS = 256;
template <int R>
__global__ void __launch_bounds__(RP_THREADS) rowpanel128_kernel(const double* __restrict__ A, long long lda,
                                                                 const double* __restrict__ Bm, long long ldb,
                                                                 const double* C, long long ldc, double* out, long long ldo,
                                                                 int M, double alpha, double beta) {
    extern __shared__ __align__(16) double rp_smem[];
    double* Bs = rp_smem;                      // [128][RP_LD]
    double* As = rp_smem + 128 * RP_LD;        // [R][RP_LD]
    constexpr int MI = R / 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t4 = lane & 3;
    const int r0 = blockIdx.x * R;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {          // 128 rows x 16 chunks of 16 bytes
            const int c = tid + i * RP_THREADS, row = c >> 4, kc = (c & 15) * 2 + 32 * g;
            cp_async_zfill<16>(Bs + row * RP_LD + kc, Bm + (long long)row * ldb + kc, 16);
        }
#pragma unroll
        for (int i = 0; i < (R * 16 + RP_THREADS - 1) / RP_THREADS; ++i) {
            const int c = tid + i * RP_THREADS;
            if (c < R * 16) {
                const int row = c >> 4, kc = (c & 15) * 2 + 32 * g, gr = r0 + row;
                cp_async_zfill<16>(As + row * RP_LD + kc, gr < M ? A + (long long)gr * lda + kc : A, gr < M ? 16 : 0);
            }
        }
        cp_async_commit();
    }
    const int n0 = warp * 16;
    double2 cin[MI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            const int r = r0 + mi * 8 + gq;
            cin[mi][ni] = (beta != 0.0 && r < M) ? *reinterpret_cast<const double2*>(C + (long long)r * ldc + n0 + ni * 8 + 2 * t4)
                                                 : make_double2(0.0, 0.0);
        }
    double acc[MI][2][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const double* Ap = As + gq * RP_LD + t4;
    const double* Bp = Bs + (n0 + gq) * RP_LD + t4;
    auto chunk = [&](int g) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const int k = 32 * g + 4 * k4;
            double a[MI], b[2];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) a[mi] = Ap[mi * 8 * RP_LD + k];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) b[ni] = Bp[ni * 8 * RP_LD + k];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
        }
    };
    cp_async_wait<3>(); __syncthreads(); chunk(0);
    cp_async_wait<2>(); __syncthreads(); chunk(1);
    cp_async_wait<1>(); __syncthreads(); chunk(2);
    cp_async_wait<0>(); __syncthreads(); chunk(3);
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
            const int r = r0 + mi * 8 + gq;
            if (r < M)
                *reinterpret_cast<double2*>(out + (long long)r * ldo + n0 + ni * 8 + 2 * t4) =
                    make_double2(fma(alpha, acc[mi][ni][0], beta * cin[mi][ni].x), fma(alpha, acc[mi][ni][1], beta * cin[mi][ni].y));
        }
}

// returns 1 when the specialised kernel was launched, 0 when the operands do not qualify (caller falls back), < 0 on error
static int launch_rowpanel128(const double* A, long long lda, const double* Bm, long long ldb, const double* C, long long ldc,
                              double* out, long long ldo, int M, double alpha, double beta, cudaStream_t st) {
    if (g_tuning[8] == 1) return 0;                       // tuning key 8 = 1: general GEMM kernel (comparison)
    if (M <= 0) return 1;
    if (!(aligned16(A) && aligned16(Bm) && aligned16(out) && (C == nullptr || aligned16(C)) && lda % 2 == 0 && ldb % 2 == 0 &&
          ldc % 2 == 0 && ldo % 2 == 0))
        return 0;
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] {
        err = cudaFuncSetAttribute(rowpanel128_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + 32) * RP_LD * 8);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(rowpanel128_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + 16) * RP_LD * 8);
    });
    PPBO_CUDA_CHECK(err);
    if (C == nullptr) { C = out; ldc = ldo; }
    if (M > PPBO_SM_COUNT * 16)
        PPBO_CL rowpanel128_kernel<32><<<ceil_div(M, 32), RP_THREADS, (128 + 32) * RP_LD * 8, st>>>(A, lda, Bm, ldb, C, ldc, out, ldo, M, alpha, beta);
    else
        PPBO_CL rowpanel128_kernel<16><<<ceil_div(M, 16), RP_THREADS, (128 + 16) * RP_LD * 8, st>>>(A, lda, Bm, ldb, C, ldc, out, ldo, M, alpha, beta);
    PPBO_LAUNCH_CHECK();
    return 1;
}

// ------------------------------------------------------------------------------------------- rank-128 update of the trailing matrix
// C[lower tiles] -= P P^T with P = [n x 128] (row stride ldp), C = [n x n] (row stride ldc): the bulk of every Cholesky step.
// The general kernel runs this shape at ~17-20 TFLOP/s (55 % of the DMMA peak): with K = 128 a 128 x 64 tile moves 320 KB
// through L2 for 2.1 MFLOP, its 3-stage ring is barely filled before the 8 k-tiles are over, and C is only requested in the
// epilogue.  Here a CTA takes a 128 x 64 tile, brings both operand panels completely into shared memory (four 32-column
// cp.async groups, compute starts on the first), requests its C fragments up front and gives every warp a 32 x 32 register tile
// (8 fragment loads per 16 DMMA).  Tiles that touch the lower triangle only; a tile on the diagonal is computed in full (the
// strictly upper part of the matrix is never read).
constexpr int SY_BM = 128, SY_BN = 64, SY_THREADS = 256;
constexpr int SY_SMEM_BYTES = (SY_BM + SY_BN) * RP_LD * 8;
__global__ void __launch_bounds__(SY_THREADS) syrk128_kernel(const double* __restrict__ P, long long ldp, double* __restrict__ C,
                                                             long long ldc, int n) {
    extern __shared__ __align__(16) double sy_smem[];
    double* As = sy_smem;                      // [128][RP_LD]  rows of the row tile
    double* Bs = sy_smem + SY_BM * RP_LD;      // [64][RP_LD]   rows of the column tile
    // tile (ti, tj): row tile ti owns column tiles 0 .. 2 ti + 1
    const long long L = blockIdx.x;
    int ti = (int)((sqrt(1.0 + 4.0 * (double)L) - 1.0) * 0.5);
    while ((long long)(ti + 1) * (ti + 2) <= L) ++ti;
    while ((long long)ti * (ti + 1) > L) --ti;
    const int tj = (int)(L - (long long)ti * (ti + 1));
    const int m0 = ti * SY_BM, n0 = tj * SY_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int i = 0; i < SY_BM * 16 / SY_THREADS; ++i) {
            const int c = tid + i * SY_THREADS, row = c >> 4, kc = (c & 15) * 2 + 32 * g, gr = m0 + row;
            cp_async_zfill<16>(As + row * RP_LD + kc, gr < n ? P + (long long)gr * ldp + kc : P, gr < n ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < SY_BN * 16 / SY_THREADS; ++i) {
            const int c = tid + i * SY_THREADS, row = c >> 4, kc = (c & 15) * 2 + 32 * g, gr = n0 + row;
            cp_async_zfill<16>(Bs + row * RP_LD + kc, gr < n ? P + (long long)gr * ldp + kc : P, gr < n ? 16 : 0);
        }
        cp_async_commit();
    }
    const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 32;       // 4 x 2 warps, 32 x 32 each
    double2 cin[4][4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int r = m0 + wm0 + mi * 8 + gq, c = n0 + wn0 + ni * 8 + 2 * t4;
            cin[mi][ni] = (r < n && c + 1 < n) ? *reinterpret_cast<const double2*>(C + (long long)r * ldc + c)
                                               : make_double2((r < n && c < n) ? C[(long long)r * ldc + c] : 0.0, 0.0);
        }
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const double* Ap = As + (wm0 + gq) * RP_LD + t4;
    const double* Bp = Bs + (wn0 + gq) * RP_LD + t4;
    auto chunk = [&](int g) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const int k = 32 * g + 4 * k4;
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = Ap[mi * 8 * RP_LD + k];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bp[ni * 8 * RP_LD + k];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
        }
    };
    cp_async_wait<3>(); __syncthreads(); chunk(0);
    cp_async_wait<2>(); __syncthreads(); chunk(1);
    cp_async_wait<1>(); __syncthreads(); chunk(2);
    cp_async_wait<0>(); __syncthreads(); chunk(3);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int r = m0 + wm0 + mi * 8 + gq, c = n0 + wn0 + ni * 8 + 2 * t4;
            if (r >= n) continue;
            double* o = C + (long long)r * ldc + c;
            if (c + 1 < n) *reinterpret_cast<double2*>(o) = make_double2(cin[mi][ni].x - acc[mi][ni][0], cin[mi][ni].y - acc[mi][ni][1]);
            else if (c < n) o[0] = cin[mi][ni].x - acc[mi][ni][0];
        }
}

// Strip version: ncu showed the tile kernel above with the tensor pipe active only 58 % of the time -- 31k clocks per 128 x 64 tile for
// 16.4k clocks of DMMA, because a CTA that owns one tile (and the whole SM: 203 KB) cannot overlap its operand loads, its C
// read-modify-write and its launch with anything.  Here a CTA keeps the 128-row operand panel resident and walks along a strip of up
// to SY_STRIP column tiles of 32: the next 32 x 128 column panel (cp.async, double buffered) and the next C fragments are requested
// while the current tile is on the tensor pipe, so the prologue is paid once per strip.
constexpr int SYS_BN = 32, SY_STRIP = 8;
constexpr int SYS_SMEM_BYTES = (SY_BM + 2 * SYS_BN) * RP_LD * 8;
__global__ void __launch_bounds__(SY_THREADS) syrk128_strip_kernel(const double* __restrict__ P, long long ldp, double* __restrict__ C,
                                                                   long long ldc, int n, int strip) {
    extern __shared__ __align__(16) double sy_smem[];
    double* As = sy_smem;                      // [128][RP_LD]
    double* Bs = sy_smem + SY_BM * RP_LD;      // 2 x [32][RP_LD]
    const int ti = gridDim.y - 1 - blockIdx.y, m0 = ti * SY_BM;    // long rows first: the tail of the launch is made of short strips
    const int ncols = min(m0 + SY_BM, n);                          // columns 0 .. ncols-1 touch the lower triangle of this row tile
    const int tiles = (ncols + SYS_BN - 1) / SYS_BN;
    const int t_begin = blockIdx.x * strip, t_end = min(t_begin + strip, tiles);
    if (t_begin >= tiles) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t4 = lane & 3;
    auto load_b = [&](int t, int buf) {                             // column panel of tile t: rows 32 t .. 32 t + 31 of P
        double* dst = Bs + buf * SYS_BN * RP_LD;
#pragma unroll
        for (int i = 0; i < SYS_BN * 64 / SY_THREADS; ++i) {        // 32 rows x 64 chunks of 16 bytes
            const int c = tid + i * SY_THREADS, row = c >> 6, kc = (c & 63) * 2, gr = t * SYS_BN + row;
            cp_async_zfill<16>(dst + row * RP_LD + kc, gr < n ? P + (long long)gr * ldp + kc : P, gr < n ? 16 : 0);
        }
    };
#pragma unroll
    for (int i = 0; i < SY_BM * 64 / SY_THREADS; ++i) {
        const int c = tid + i * SY_THREADS, row = c >> 6, kc = (c & 63) * 2, gr = m0 + row;
        cp_async_zfill<16>(As + row * RP_LD + kc, gr < n ? P + (long long)gr * ldp + kc : P, gr < n ? 16 : 0);
    }
    load_b(t_begin, 0);
    cp_async_commit();
    const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 16;       // 4 x 2 warps, 32 x 16 each
    auto load_c = [&](int t, double2 (&cin)[4][2]) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int r = m0 + wm0 + mi * 8 + gq, c = t * SYS_BN + wn0 + ni * 8 + 2 * t4;
                cin[mi][ni] = (r < n && c + 1 < n) ? *reinterpret_cast<const double2*>(C + (long long)r * ldc + c)
                                                   : make_double2((r < n && c < n) ? C[(long long)r * ldc + c] : 0.0, 0.0);
            }
    };
    double2 cin[4][2], cnext[4][2];
    load_c(t_begin, cin);
    const double* Ap = As + (wm0 + gq) * RP_LD + t4;
    for (int t = t_begin; t < t_end; ++t) {
        const int buf = (t - t_begin) & 1;
        cp_async_wait<0>();
        __syncthreads();                                            // panel of tile t has landed; everyone is done with tile t-1
        if (t + 1 < t_end) {
            load_b(t + 1, buf ^ 1);
            load_c(t + 1, cnext);
        }
        cp_async_commit();
        double acc[4][2][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const double* Bp = Bs + buf * SYS_BN * RP_LD + (wn0 + gq) * RP_LD + t4;
#pragma unroll 8
        for (int k = 0; k < 128; k += 4) {
            double a[4], b[2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = Ap[mi * 8 * RP_LD + k];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) b[ni] = Bp[ni * 8 * RP_LD + k];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int r = m0 + wm0 + mi * 8 + gq, c = t * SYS_BN + wn0 + ni * 8 + 2 * t4;
                if (r < n) {
                    double* o = C + (long long)r * ldc + c;
                    if (c + 1 < n) *reinterpret_cast<double2*>(o) = make_double2(cin[mi][ni].x - acc[mi][ni][0], cin[mi][ni].y - acc[mi][ni][1]);
                    else if (c < n) o[0] = cin[mi][ni].x - acc[mi][ni][0];
                }
                cin[mi][ni] = cnext[mi][ni];
            }
    }
}

// 1: launched, 0: operands do not qualify (caller falls back to the general kernel), < 0: error
static int launch_syrk128(const double* P, long long ldp, double* C, long long ldc, int n, cudaStream_t st) {
    if (g_tuning[9] == 1) return 0;                       // tuning key 9 = 1: general GEMM kernel; 2: one tile per CTA (comparison)
    if (n <= 0) return 1;
    if (!(aligned16(P) && aligned16(C) && ldp % 2 == 0 && ldc % 2 == 0)) return 0;
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] {
        err = cudaFuncSetAttribute(syrk128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_BYTES);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(syrk128_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SYS_SMEM_BYTES);
    });
    PPBO_CUDA_CHECK(err);
    const int tm = ceil_div(n, SY_BM);
    if (g_tuning[9] == 2) {
        PPBO_CL syrk128_kernel<<<(unsigned)((long long)tm * (tm + 1)), SY_THREADS, SY_SMEM_BYTES, st>>>(P, ldp, C, ldc, n);
    } else {
        // strip length: one CTA per SM, so the launch takes ceil(items / SMs) rounds of (strip tiles + one panel load); pick the
        // length that minimises that (ncu: 26 % of the SM time was idle tail with a fixed length of 8 at n = 4480)
        int strip = SY_STRIP;
        double best = 1e300;
        for (int sl = 3; sl <= 12; ++sl) {
            long long items = 0;
            for (int t = 0; t < tm; ++t) items += ceil_div(ceil_div(min((t + 1) * SY_BM, n), SYS_BN), sl);
            const double cost = (double)ceil_div_ll(items, PPBO_SM_COUNT) * (sl * 8.2 + 6.0);
            if (cost < best) { best = cost; strip = sl; }
        }
        if (g_tuning[10] > 0) strip = g_tuning[10];
        const int max_strips = ceil_div(ceil_div(min(tm * SY_BM, n), SYS_BN), strip);
        PPBO_CL syrk128_strip_kernel<<<dim3(max_strips, tm), SY_THREADS, SYS_SMEM_BYTES, st>>>(P, ldp, C, ldc, n, strip);
    }
    PPBO_LAUNCH_CHECK();
    return 1;
}

// ------------------------------------------------------------------------------------------- GEMV
// y = A x, row-major A: one warp per row, 16-byte loads when aligned.  HBM-bound (8 M N bytes).
// skip (here and in the block-inverse solve kernels): optional device flag; a non-zero value turns the launch into a no-op.  The
// chord steps of the Newton iterations are queued in batches, and the steps behind the one that converged must cost nothing.
__global__ void __launch_bounds__(256) gemv_kernel(const double* __restrict__ A, long long lda, int M, int N,
                                                   const double* __restrict__ x, double* __restrict__ y, int vec,
                                                   const double* __restrict__ skip) {
    if (skip && *skip != 0.0) return;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int row = warp; row < M; row += nwarps) {
        const double* a = A + (long long)row * lda;
        double s0 = 0.0, s1 = 0.0;
        if (vec) {
            const double2* a2 = reinterpret_cast<const double2*>(a);
            const double2* x2 = reinterpret_cast<const double2*>(x);
            const int n2 = N >> 1;
            for (int j = lane; j < n2; j += 32) {
                const double2 av = a2[j], xv = x2[j];
                s0 = fma(av.x, xv.x, s0);
                s1 = fma(av.y, xv.y, s1);
            }
            if ((N & 1) && lane == 0) s0 = fma(a[N - 1], x[N - 1], s0);
        } else {
            for (int j = lane; j < N; j += 32) s0 = fma(a[j], x[j], s0);
        }
        double s = s0 + s1;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[row] = s;
    }
}

int gemv(const double* A, long long lda, int M, int N, const double* x, double* y, cudaStream_t st, const double* skip) {
    if (M <= 0) return PPBO_OK;
    const int vec = aligned16(A) && aligned16(x) && lda % 2 == 0;
    const int blocks = min(ceil_div(M, 8), PPBO_SM_COUNT * 8);
    PPBO_CL gemv_kernel<<<blocks, 256, 0, st>>>(A, lda, M, N, x, y, vec, skip);
    PPBO_LAUNCH_CHECK();
    return PPBO_OK;
}

// ------------------------------------------------------------------------------------------- Cholesky
// Diagonal block: factor the jb x jb (jb <= 128) lower block and invert the factor, one CTA, everything in shared memory.
// dinv receives inv(L_jj) as a dense NB x NB row-major lower-triangular block (zeros above the diagonal and beyond jb).
//
// The block is processed as 4 sub-columns of 32: warp 0 factors the 32 x 32 diagonal piece and inverts it entirely in
// registers (lane i owns row i; pivots and multipliers travel by warp shuffle, no block barrier), then all 16 warps form the
// sub-panel L21 = A21 inv(L11)^T and the trailing update as small shared-memory GEMMs.  The 128 x 128 inverse is assembled from
// the four 32 x 32 inverses by two levels of  inv([[A,0],[C,B]]) = [[A^-1,0],[-B^-1 C A^-1, B^-1]].
// ~35 us per block against ~300 us for a column-by-column version with three block barriers per column.
constexpr int PAIR_MIN_REM = 2048;                    // remainders above this defer / pair their trailing updates (K = 256)
constexpr int POTF2_THREADS = 512;
// Row strides == 4 (mod 16) doubles: the 8-byte DMMA fragment loads of a half-warp (8 rows x 4 consecutive k, or 4 k-rows x 8
// columns) then hit 16 distinct bank pairs.  With the odd strides used before (129 / 65 / 33) they were 3-4-way conflicted and the
// tensor-pipe phases of the diagonal-block kernel were shared-memory bound (80.6k -> 70.5k clocks per block); lane-per-row accesses
// (the warp Cholesky's 32 loads and stores, the row-wise sub-panel solve) are 4-way conflicted instead.
// Tried and not kept (measured): sub-panel on the tensor pipe after the inversion (73.0k: the extra accumulators push the
// kernel over its 128-register budget and the inversion spills), block rows of the inverse finished next to the inversion instead
// of next to the pivot chain (85.4k).
constexpr int POTF2_LDS = CHOL_NB + 4;                 // row stride of the 128 x 128 work matrix
constexpr int POTF2_SUB = 32;
constexpr int POTF2_LDR = POTF2_SUB + 4;               // row stride of the 32 x 32 inverse blocks
constexpr int POTF2_LDT = 64 + 4;                      // row stride of the scratch matrix (up to 96 x 32 or 64 x 64)
constexpr int POTF2_SMEM_DOUBLES = CHOL_NB * POTF2_LDS + 4 * POTF2_SUB * POTF2_LDR + 96 * POTF2_LDT;

// C[i][j] = (acc ? C[i][j] : 0) + sign * sum_k A[i][k] * (B_KMAJOR ? B[k][j] : B[j][k]),  i < m, j < n, all in shared memory.
// lower_only: only i >= j + diag_shift is computed.  Thread t owns rows {ti + r * tiles_m} x columns {tj + c * tiles_n} so that
// the lanes of a warp read consecutive rows (conflict-free with the odd row strides used here).
template <bool B_KMAJOR>
__device__ __forceinline__ void smem_gemm(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int m, int n,
                                          int K, double sign, bool acc, bool lower_only) {
    const int tiles_m = (m + 1) >> 1, tiles_n = (n + 1) >> 1;
    for (int t = threadIdx.x; t < tiles_m * tiles_n; t += POTF2_THREADS) {
        const int ti = t % tiles_m, tj = t / tiles_m;
        const int i0 = ti, i1 = ti + tiles_m, j0 = tj, j1 = tj + tiles_n;
        if (lower_only && i1 < j0 && i0 < j0) continue;
        const bool vi1 = i1 < m, vj1 = j1 < n;
        const double* a0 = A + i0 * lda;
        const double* a1 = A + (vi1 ? i1 : i0) * lda;
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
        if (B_KMAJOR) {
            const int jj1 = vj1 ? j1 : j0;
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const double x0 = a0[k], x1 = a1[k], y0 = B[k * ldb + j0], y1 = B[k * ldb + jj1];
                c00 = fma(x0, y0, c00); c01 = fma(x0, y1, c01); c10 = fma(x1, y0, c10); c11 = fma(x1, y1, c11);
            }
        } else {
            const double* b0 = B + j0 * ldb;
            const double* b1 = B + (vj1 ? j1 : j0) * ldb;
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                const double x0 = a0[k], x1 = a1[k], y0 = b0[k], y1 = b1[k];
                c00 = fma(x0, y0, c00); c01 = fma(x0, y1, c01); c10 = fma(x1, y0, c10); c11 = fma(x1, y1, c11);
            }
        }
        auto put = [&](int i, int j, double v) {
            if (lower_only && i < j) return;
            double* c = C + i * ldc + j;
            *c = (acc ? *c : 0.0) + sign * v;
        };
        put(i0, j0, c00);
        if (vj1) put(i0, j1, c01);
        if (vi1) put(i1, j0, c10);
        if (vi1 && vj1) put(i1, j1, c11);
    }
}

// Same contract as smem_gemm, on the FP64 tensor pipe: each warp owns 16 x 16 output tiles (2 x 2 DMMA.8x8x4 tiles) and
// loads one double per lane and fragment, i.e. 1/16 shared-memory load per FMA instead of 1 (the scalar version is
// LDS-bandwidth bound and took ~33 us of the diagonal-block kernel).  m, n multiples of 16, K a multiple of 4.
// KR restricts the k range of a tile when one operand is triangular (one SM delivers only ~127 FP64 tensor flop per clock, so
// skipping the structural zeros is worth it): KR_GE_J: B[k][j] = 0 for k < j;  KR_LE_I: A[i][k] = 0 for k > i.
// Warps [warp0, warp0 + nwarps) take part.
enum { KR_FULL = 0, KR_GE_J = 1, KR_LE_I = 2 };
// smem_gemm_mma_w: the calling warp is participant `warp` of `nwarps` (warp < 0: not taking part).
template <bool B_KMAJOR, int KR = KR_FULL>
__device__ __forceinline__ void smem_gemm_mma_w(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int m, int n,
                                                int K, double sign, bool acc, bool lower_only, int warp, int nwarps) {
    const int lane = threadIdx.x & 31, gq = lane >> 2, t4 = lane & 3;
    if (warp < 0 || warp >= nwarps) return;
    const int tm = m >> 4, tn = n >> 4;
    for (int t = warp; t < tm * tn; t += nwarps) {
        const int ti = t % tm, tj = t / tm;
        if (lower_only && ti < tj) continue;
        const int i0 = ti * 16, j0 = tj * 16;
        const int k_lo = (KR == KR_GE_J) ? j0 : 0, k_hi = (KR == KR_LE_I) ? min(K, i0 + 16) : K;
        double c[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
        const double* a0 = A + (i0 + gq) * lda + t4;
        const double* a1 = a0 + 8 * lda;
#pragma unroll 4
        for (int k = k_lo; k < k_hi; k += 4) {
            const double x0 = a0[k], x1 = a1[k];
            double y0, y1;
            if (B_KMAJOR) {
                y0 = B[(k + t4) * ldb + j0 + gq];
                y1 = B[(k + t4) * ldb + j0 + 8 + gq];
            } else {
                y0 = B[(j0 + gq) * ldb + k + t4];
                y1 = B[(j0 + 8 + gq) * ldb + k + t4];
            }
            dmma884(c[0][0], x0, y0);
            dmma884(c[0][1], x0, y1);
            dmma884(c[1][0], x1, y0);
            dmma884(c[1][1], x1, y1);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = i0 + mi * 8 + gq, j = j0 + ni * 8 + 2 * t4 + e;
                    if (lower_only && i < j) continue;
                    double* dst = C + i * ldc + j;
                    *dst = (acc ? *dst : 0.0) + sign * c[mi][ni][e];
                }
    }
}

template <bool B_KMAJOR, int KR = KR_FULL>
__device__ __forceinline__ void smem_gemm_mma(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int m, int n,
                                              int K, double sign, bool acc, bool lower_only, int warp0 = 0,
                                              int nwarps = 16) {
    smem_gemm_mma_w<B_KMAJOR, KR>(C, ldc, A, lda, B, ldb, m, n, K, sign, acc, lower_only, (int)(threadIdx.x >> 5) - warp0, nwarps);
}

// Helper warps of the diagonal-block kernel: the 12 warps that do NOT share a scheduler / FP64 pipe with warp 0 (warp w runs on
// sub-partition w & 3).  Work that runs concurrently with warp 0's pivot chain is confined to them: with all 15 other warps the
// chain of the last piece slowed from 3.7k to 8.6k clocks.
constexpr int POTF2_HELPERS = 12;
__device__ __forceinline__ int potf2_helper_index() {
    const int w = threadIdx.x >> 5;
    return (w & 3) ? (w >> 2) * 3 + (w & 3) - 1 : -1;
}

// Branch-free FP64 reciprocal / reciprocal square root for the pivot chain: MUFU seed (>= 20 bits) + the Newton sequence CUDA's
// own division / rsqrt use, WITHOUT their slow-path branches (denormal / huge operands cannot occur for a pivot that passes the
// d > 0 test; a failed pivot is reported, its garbage is never used).  The library versions put a BRA / CALL into every column of
// the unrolled factorisation, which splits it into basic blocks and keeps ptxas from overlapping the pivot chain of column k+1
// with the rank-1 update of column k (measured: 380-680 clocks per column instead of the ~150 of the dependency chain).
__device__ __forceinline__ double rcp_nr(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    e = fma(e, e, e);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}
__device__ __forceinline__ double rsqrt_nr(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d * y, y, 1.0);                       // 1 - d y^2
    y = fma(y * e, fma(e, 0.375, 0.5), y);                // y (1 + e/2 + 3 e^2 / 8)
    e = fma(-d * y, y, 1.0);
    return fma(y * e, 0.5, y);
}

// warp-level Cholesky of a 32 x 32 block held in shared memory (row stride lds), in place, UNSCALED: on return the block holds
// r[i][k] = L[i][k] L[k][k] (lower triangle) and idiag[k] = 1 / L[k][k]; the caller scales column k by idiag[k] (all threads, after
// its barrier).  Lane i owns row i in registers; column k of the running Schur complement is exchanged through a ring of eight
// shared-memory lines read back as broadcast loads (one __syncwarp per column).
// The dependency chain of a column is  LDS pivot -> reciprocal -> one FMA -> STS -> __syncwarp.  To keep that chain free of other
// work whatever ptxas decides to do (measured: 116 to 294 clocks per column for the same source in three builds, depending on
// how the rank-1 updates of the other columns were interleaved), the columns are processed in blocks of four with DELAYED
// updates: inside a block only the block's own columns are updated per step (<= 2 FMAs next to the chain), and the rank-4 update
// of all later columns follows as one throughput-bound stretch.  Finished columns leave the registers at once.
// wsm: [0,256) exchange lines, [256,288) 1 / L[k][k], [288,320) pivots.  Returns 0 or the 1-based index of the first non-positive
// (or NaN) pivot, same value in every lane.
constexpr int POTF2_WSM = 320, POTF2_IDIAG = 256, POTF2_PIV = 288;
__device__ __forceinline__ int warp_potrf32(double* Sb, int lds, double* wsm, long long* tdbg = nullptr) {
    const int lane = threadIdx.x & 31;
    if (tdbg && lane == 0) tdbg[0] = clock64();
    double* idiag = wsm + POTF2_IDIAG;
    double* piv = wsm + POTF2_PIV;
    double r[POTF2_SUB];
#pragma unroll
    for (int j = 0; j < POTF2_SUB; ++j) r[j] = (j <= lane) ? Sb[lane * lds + j] : 0.0;
    wsm[lane] = r[0];
    __syncwarp();
    if (tdbg && lane == 0) tdbg[1] = clock64();
#pragma unroll
    for (int b = 0; b < POTF2_SUB / 4; ++b) {
        double mine[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = 4 * b + kk;
            const double* c = wsm + (k & 7) * 32;
            const double d = c[k];
            const double inv = rcp_nr(d);
            if (k + 1 < POTF2_SUB) {
                if (kk == 3) {               // first column of the next block: the three pending updates do not wait for `inv`
#pragma unroll
                    for (int q = 0; q < 3; ++q) r[k + 1] = fma(-mine[q], wsm[((4 * b + q) & 7) * 32 + k + 1], r[k + 1]);
                }
                const double t = r[k] * c[k + 1];
                r[k + 1] = fma(-t, inv, r[k + 1]);
                wsm[((k + 1) & 7) * 32 + lane] = r[k + 1];      // lanes <= k hold 0 there; slot last read 8 columns ago
                __syncwarp();
            }
            mine[kk] = r[k] * inv;
#pragma unroll
            for (int j = k + 2; j < 4 * b + 4; ++j) r[j] = fma(-mine[kk], c[j], r[j]);
            if (lane == 0) piv[k] = d;
            if (k <= lane) Sb[lane * lds + k] = r[k];           // final (unscaled)
        }
        const double* c0 = wsm + ((4 * b) & 7) * 32;            // slots 4b .. 4b+3 of the ring are consecutive
#pragma unroll
        for (int j = 4 * b + 5; j < POTF2_SUB; ++j)
            r[j] = fma(-mine[3], c0[96 + j], fma(-mine[2], c0[64 + j], fma(-mine[1], c0[32 + j], fma(-mine[0], c0[j], r[j]))));
    }
    __syncwarp();
    if (tdbg && lane == 0) tdbg[2] = clock64();
    const double dk = piv[lane];
    const bool ok = dk > 0.0;                                          // false for NaN
    const unsigned badmask = __ballot_sync(0xffffffffu, !ok);
    idiag[lane] = rsqrt_nr(ok ? dk : 1.0);
    __syncwarp();
    if (tdbg && lane == 0) tdbg[3] = clock64();
    return badmask ? __ffs(badmask) : 0;
}

// all threads: L11[i][j] = r[i][j] / L[j][j] on the lower triangle of the 32 x 32 piece warp_potrf32 left unscaled
__device__ __forceinline__ void potf2_scale_piece(double* Sb, int lds, const double* idiag) {
    for (int e = threadIdx.x; e < POTF2_SUB * POTF2_SUB; e += POTF2_THREADS) {
        const int i = e >> 5, j = e & 31;
        if (j <= i) Sb[i * lds + j] *= idiag[j];
    }
}

// inverse of the 32 x 32 lower-triangular block Sb (one warp): lane c owns column c of R = L^-1,
// R[i][c] = ((i == c) - sum_{c <= p < i} L[i][p] R[p][c]) / L[i][i];  Rb row stride POTF2_LDR
__device__ __forceinline__ void warp_inv32(const double* Sb, int lds, double* Rb, const double* idiag) {
    const int lane = threadIdx.x & 31;
    double x[POTF2_SUB];
#pragma unroll
    for (int i = 0; i < POTF2_SUB; ++i) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int p = 0; p + 1 < i; p += 2) {
            s0 = fma(Sb[i * lds + p], x[p], s0);
            s1 = fma(Sb[i * lds + p + 1], x[p + 1], s1);
        }
        if (i & 1) s0 = fma(Sb[i * lds + i - 1], x[i - 1], s0);
        x[i] = (i < lane) ? 0.0 : (((i == lane) ? 1.0 : 0.0) - (s0 + s1)) * idiag[i];
    }
#pragma unroll
    for (int j = 0; j < POTF2_SUB; ++j) Rb[j * POTF2_LDR + lane] = x[j];          // row j, column lane
}

// one row of the sub-panel: x = a . inv(L11)^T by forward substitution in registers (column-oriented: the dependency chain is
// one multiply + one FMA per column); L11 entries are broadcast shared-memory loads.  Writes the row back and into the scratch
// matrix that feeds the trailing update.
__device__ __forceinline__ void row_trsv32(double* arow, const double* L11, int lds, const double* idiag, double* trow) {
    double a[POTF2_SUB];
#pragma unroll
    for (int j = 0; j < POTF2_SUB; ++j) a[j] = arow[j];
#pragma unroll
    for (int j = 0; j < POTF2_SUB; ++j) {
        const double x = a[j] * idiag[j];
        a[j] = x;
#pragma unroll
        for (int q = j + 1; q < POTF2_SUB; ++q) a[q] = fma(-x, L11[q * lds + j], a[q]);
    }
#pragma unroll
    for (int j = 0; j < POTF2_SUB; ++j) {
        arow[j] = a[j];
        trow[j] = a[j];
    }
}

__device__ __forceinline__ void potf2_named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Diagonal-block kernel.  Per 32-column sub-step sb (c0 = 32 sb):
//   phase 1: warp 0 factors the 32 x 32 diagonal piece  ||  warps 1..15 finish block row sb-1 (see below)
//   phase 2: warp 0 inverts the piece (R_sb)            ||  warps 1..15 solve the sub-panel row by row and apply the trailing update
// Block row i of the factor (L_i0 .. L_ii) is final once its diagonal piece is factored, and R_i exists after phase 2 of sub-step i.
// "Finishing" block row i (finish_block_row) is everything that used to follow the factorisation as a serial tail:
//   write L's block row back to A;  X_i,0:i-1 = -R_i (L_i,0:i-1 X_0:i-1,0:i-1)  (block row i of X = inv(L), by forward substitution
//   over the earlier block rows of X, which overwrite L in S);  X_ii = R_i;  store the block row of X to dinv.
// Rows 0..2 are finished while warp 0 factors the next piece (the other 15 warps would be idle), so only row 3 remains after
// the last piece: the old tail (two levels of block inversion + a 128 KB store: 22k clocks) shrinks to ~5k.
// dinv must be zero above the diagonal and beyond jb on entry (potrf_lower clears it): only the lower triangle is stored.
__device__ __forceinline__ void potf2_helper_barrier() { potf2_named_barrier(2, POTF2_HELPERS * 32); }

// T = L_i,0:i-1 . X_0:i-1,0:i-1 -> Tm (row stride ldt); helper warps, no barrier inside
__device__ __forceinline__ void potf2_row_product(const double* S, double* Tm, int ldt, int i) {
    smem_gemm_mma_w<true, KR_GE_J>(Tm, ldt, S + (32 * i) * POTF2_LDS, POTF2_LDS, S, POTF2_LDS, 32, 32 * i, 32 * i, 1.0, false, false,
                                   potf2_helper_index(), POTF2_HELPERS);
}

// lower triangle of rows [r0, r0 + 32) of S (columns 0 .. row) -> global dst (row stride ldd), rows < jb only; the calling
// thread is participant t of nthr
__device__ __forceinline__ void potf2_store_rows(const double* S, double* __restrict__ dst, long long ldd, int r0, int jb, int t,
                                                 int nthr) {
    const int ncol = r0 + 32;
    for (int e = t; e < 32 * ncol; e += nthr) {
        const int i = r0 + e / ncol, j = e % ncol;
        if (i < jb && j <= i) dst[(long long)i * ldd + j] = S[i * POTF2_LDS + j];
    }
}

// helper warps: finish block row i (0 <= i <= 2) while warp 0 factors piece i + 1.  Tm is free in phase 1.
__device__ __forceinline__ void potf2_finish_block_row(double* S, const double* Rd, double* Tm, int i, double* __restrict__ A,
                                                       long long lda, double* __restrict__ dinv, int jb) {
    const int hw = potf2_helper_index();
    if (hw < 0) return;
    const int t0 = hw * 32 + (threadIdx.x & 31), nthr = POTF2_HELPERS * 32, r0 = 32 * i;
    potf2_store_rows(S, A, lda, r0, jb, t0, nthr);                              // factor rows -> A
    if (i > 0) potf2_row_product(S, Tm, POTF2_LDT, i);
    potf2_helper_barrier();                                                      // L_i* no longer needed in S
    const double* Ri = Rd + i * POTF2_SUB * POTF2_LDR;
    if (i > 0)                                                                   // X_i,0:i-1 = -R_i T   (R_i[r][k] = 0 for k > r)
        smem_gemm_mma_w<true, KR_LE_I>(S + r0 * POTF2_LDS, POTF2_LDS, Ri, POTF2_LDR, Tm, POTF2_LDT, 32, r0, 32, -1.0, false, false, hw,
                                       POTF2_HELPERS);
    for (int e = t0; e < POTF2_SUB * POTF2_SUB; e += nthr) {                     // X_ii = R_i (zeros above the diagonal)
        const int r = e >> 5, c = e & 31;
        S[(r0 + r) * POTF2_LDS + r0 + c] = Ri[r * POTF2_LDR + c];
    }
    potf2_helper_barrier();
    potf2_store_rows(S, dinv, CHOL_NB, r0, jb, t0, nthr);                        // inverse rows -> dinv
}

__global__ void __launch_bounds__(POTF2_THREADS) potf2_inv_kernel(double* __restrict__ A, long long lda, int jb,
                                                                  double* __restrict__ dinv, int* __restrict__ info,
                                                                  int block_offset, long long* __restrict__ dbg) {
    extern __shared__ double sm[];
#ifdef PPBO_PDL
    asm volatile("griddepcontrol.launch_dependents;");       // see gemm_nt_store_kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    int dbg_i = 0;
#define POTF2_STAMP() do { if (dbg && threadIdx.x == 0) dbg[dbg_i++] = clock64(); } while (0)
    POTF2_STAMP();
    double* S = sm;                                         // 128 x 128 work matrix (lower), identity-padded beyond jb
    double* Rd = S + CHOL_NB * POTF2_LDS;                   // 4 inverted 32 x 32 diagonal pieces
    double* Tm = Rd + 4 * POTF2_SUB * POTF2_LDR;            // scratch
    __shared__ int bad_s;
    __shared__ double wsm[POTF2_WSM];
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) bad_s = 0;
#pragma unroll 8
    for (int q = 0; q < CHOL_NB * CHOL_NB / POTF2_THREADS; ++q) {
        const int e = tid + q * POTF2_THREADS;
        const int i = e >> 7, j = e & 127;
        double v = 0.0;
        if (i < jb && j <= i) v = A[(long long)i * lda + j];
        else if (i >= jb && i == j) v = 1.0;
        S[i * POTF2_LDS + j] = v;
    }
    __syncthreads();
    POTF2_STAMP();
    constexpr int LDT3 = 100;                               // row stride of the 32 x 96 product of the last block row
    for (int sb = 0; sb < 4; ++sb) {
        const int c0 = sb * POTF2_SUB, c1 = c0 + POTF2_SUB, rem = CHOL_NB - c1;
        double* S11 = S + c0 * POTF2_LDS + c0;
        if (warp == 0) {
            const int bad = warp_potrf32(S11, POTF2_LDS, wsm, dbg ? dbg + 16 + 4 * sb : nullptr);
            if (bad && tid == 0 && !bad_s) bad_s = c0 + bad;
        } else if (sb > 0) {
            potf2_finish_block_row(S, Rd, Tm, sb - 1, A, lda, dinv, jb);
        }
        __syncthreads();
        if (bad_s) break;                                   // uniform
        potf2_scale_piece(S11, POTF2_LDS, wsm + POTF2_IDIAG);
        __syncthreads();
        POTF2_STAMP();
        if (warp == 0) {
            warp_inv32(S11, POTF2_LDS, Rd + sb * POTF2_SUB * POTF2_LDR, wsm + POTF2_IDIAG);
        } else if (rem > 0) {
            const int i = tid - 32;
            if (i < rem) row_trsv32(S + (c1 + i) * POTF2_LDS + c0, S11, POTF2_LDS, wsm + POTF2_IDIAG, Tm + i * POTF2_LDT);
            potf2_named_barrier(1, POTF2_THREADS - 32);
            // trailing: S22 -= L21 L21^T (lower), warps 1..15
            smem_gemm_mma<false>(S + c1 * POTF2_LDS + c1, POTF2_LDS, Tm, POTF2_LDT, Tm, POTF2_LDT, rem, rem, POTF2_SUB, -1.0, true, true,
                                 1, POTF2_THREADS / 32 - 1);
        } else {
            // last sub-step: the factor is complete.  While warp 0 inverts the last piece: write the last block row back and
            // form T = L_3,0:2 X_0:2,0:2 (it does not need R_3)
            const int hw = potf2_helper_index();
            if (hw >= 0) potf2_store_rows(S, A, lda, c0, jb, hw * 32 + (tid & 31), POTF2_HELPERS * 32);
            potf2_row_product(S, Tm, LDT3, 3);
        }
        __syncthreads();
        POTF2_STAMP();
    }
    if (bad_s) {
        if (tid == 0) atomicCAS(info, 0, block_offset + bad_s);
        return;
    }
    POTF2_STAMP();
    // ---- tail: X_3,0:2 = -R_3 T (all 16 warps), then the last block row of the inverse goes to dinv (its diagonal piece from Rd)
    const double* R3 = Rd + 3 * POTF2_SUB * POTF2_LDR;
    smem_gemm_mma<true, KR_LE_I>(S + 96 * POTF2_LDS, POTF2_LDS, R3, POTF2_LDR, Tm, LDT3, 32, 96, 32, -1.0, false, false);
    __syncthreads();
    POTF2_STAMP();
    POTF2_STAMP();
    for (int e = tid; e < 32 * CHOL_NB; e += POTF2_THREADS) {
        const int r = e >> 7, j = e & 127, i = 96 + r;
        if (i < jb && j <= i) dinv[i * CHOL_NB + j] = (j < 96) ? S[i * POTF2_LDS + j] : R3[r * POTF2_LDR + (j - 96)];
    }
    POTF2_STAMP();
#undef POTF2_STAMP
}

long long potrf_dinv_doubles(int n) { return (long long)ceil_div(n, CHOL_NB) * CHOL_NB * CHOL_NB; }

static long long* g_potf2_dbg = nullptr;     // device buffer for the phase clocks of one diagonal-block kernel (timeline mode)

static thread_local bool g_thread_background = false;

bool thread_is_background() { return g_thread_background; }

struct CholStreams {
    cudaStream_t side = nullptr, crit = nullptr;
    cudaEvent_t panel_done[2] = {nullptr, nullptr}, rest_done[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join = nullptr;
    int init() {
        if (side) return PPBO_OK;
        // The block-column critical path (diagonal block, panel, next block column) runs on a high-priority stream and the
        // bulk of the trailing update on a low-priority one: the trailing GEMM's ~1500 short CTAs otherwise occupy every SM
        // and the one-CTA diagonal-block kernel of the next step queues behind them.
        // A background thread (ppbo_set_thread_background: the weight-space fit that runs concurrently with the GP fit) keeps
        // everything at the lowest priority; a foreground thread puts its bulk one level above that, so that its trailing updates
        // are not queued behind the background work either.
        int lo = 0, hi = 0;
        PPBO_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));        // lo: lowest priority (largest number)
        const int mid = (hi < lo) ? lo - 1 : lo;
        PPBO_CUDA_CHECK(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, g_thread_background ? lo : mid));
        PPBO_CUDA_CHECK(cudaStreamCreateWithPriority(&crit, cudaStreamNonBlocking, g_thread_background ? lo : hi));
        PPBO_CUDA_CHECK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            PPBO_CUDA_CHECK(cudaEventCreateWithFlags(&panel_done[i], cudaEventDisableTiming));
            PPBO_CUDA_CHECK(cudaEventCreateWithFlags(&rest_done[i], cudaEventDisableTiming));
        }
        PPBO_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        return PPBO_OK;
    }
};
static thread_local CholStreams g_chol;

// Right-looking blocked Cholesky with one block column of look-ahead: the next diagonal block and panel are
// factored on `st` while the bulk of the trailing update runs on a side stream.
int potrf_lower(double* A, long long lda, int n, double* dinv, int* info_d, cudaStream_t caller) {
    if (n <= 0) return PPBO_OK;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    const int potf2_smem = POTF2_SMEM_DOUBLES * (int)sizeof(double);
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, potf2_smem);
    });
    PPBO_CUDA_CHECK(attr_err);
    int rc = g_chol.init();
    if (rc) return rc;
    const int nblk = ceil_div(n, CHOL_NB);
    const bool lookahead = nblk > 3;
    cudaStream_t st = caller;
    if (lookahead && g_tuning[3] == 0) {                  // tuning key 3: 1 = keep the critical path on the caller's stream
        st = g_chol.crit;
        PPBO_CUDA_CHECK(cudaEventRecord(g_chol.fork, caller));
        PPBO_CUDA_CHECK(cudaStreamWaitEvent(st, g_chol.fork, 0));
    }
    PPBO_CUDA_CHECK(cudaMemsetAsync(info_d, 0, sizeof(int), st));
    // the diagonal-block kernel only stores the lower triangles of the inverted blocks
    PPBO_CUDA_CHECK(cudaMemsetAsync(dinv, 0, sizeof(double) * potrf_dinv_doubles(n), st));
    bool side_busy = false, pair_open = false;
    int last_rest = 0;
    // tuning key 5 = 1: record a timeline (start of every diagonal block, after panel, after look-ahead; start/end of the bulk
    // trailing updates) and print it to stderr after a device synchronise -- diagnostics only
    const bool tl = g_tuning[5] == 1 && nblk <= 64;
    // tuning key 7 = 1: programmatic dependent launch on the critical path.  Measured at n = 5000: 4.19 ms with it against
    // 4.05 ms without (the early-resident CTAs take SM slots from the trailing update), so it is off by default.
#ifdef PPBO_PDL
    const bool pdl = lookahead && g_tuning[7] == 1;
#else
    const bool pdl = false;                               // (the kernels are compiled without griddepcontrol.wait: see gemm_f64.cuh)
#endif
    static cudaEvent_t ev_d[64], ev_p[64], ev_l[64], ev_r0[64], ev_r1[64], ev_end;
    static bool ev_init = false;
    if (tl && !g_potf2_dbg) cudaMalloc(&g_potf2_dbg, 64 * sizeof(long long));
    if (tl && !ev_init) {
        for (int i = 0; i < 64; ++i) {
            cudaEventCreate(&ev_d[i]); cudaEventCreate(&ev_p[i]); cudaEventCreate(&ev_l[i]);
            cudaEventCreate(&ev_r0[i]); cudaEventCreate(&ev_r1[i]);
        }
        cudaEventCreate(&ev_end);
        ev_init = true;
    }
    bool has_l[64] = {false}, has_r[64] = {false};
    for (int b = 0; b < nblk; ++b) {
        const int j0 = b * CHOL_NB, jb = min(CHOL_NB, n - j0), j1 = j0 + jb, rem = n - j1;
        double* Ajj = A + (long long)j0 * lda + j0;
        double* dinv_b = dinv + (long long)b * CHOL_NB * CHOL_NB;
        if (tl) cudaEventRecord(ev_d[b], st);
        if (pdl) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(1);
            cfg.blockDim = dim3(POTF2_THREADS);
            cfg.dynamicSmemBytes = potf2_smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            ++g_launch_count;
            PPBO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, potf2_inv_kernel, Ajj, lda, jb, dinv_b, info_d, j0,
                                               (long long*)((tl && b == 20) ? g_potf2_dbg : nullptr)));
        } else {
            PPBO_CL potf2_inv_kernel<<<1, POTF2_THREADS, potf2_smem, st>>>(Ajj, lda, jb, dinv_b, info_d, j0,
                                                                             (tl && b == 20) ? g_potf2_dbg : nullptr);
            PPBO_LAUNCH_CHECK();
        }
        if (tl) cudaEventRecord(ev_p[b], st);
        if (rem <= 0) break;
        // panel: L21 = A21 . inv(L11)^T   (in place: each CTA owns whole rows, K == jb <= BN)
        double* A21 = A + (long long)j1 * lda + j0;
        rc = (jb == CHOL_NB && !pdl) ? launch_rowpanel128(A21, lda, dinv_b, CHOL_NB, nullptr, 0, A21, lda, rem, 1.0, 0.0, st) : 0;
        if (rc < 0) return rc;
        if (rc == 0) {
            GemmOperands g{A21, lda, 0, dinv_b, CHOL_NB, 0, rem, jb, jb};
            StoreEpilogue ep{A21, lda, 0, 1.0, 0.0, 0, 0, 1};
            g_launch_pdl = pdl;
            rc = launch_gemm_nt(g, ep, 1, st);
            g_launch_pdl = false;
            if (rc) return rc;
        }
        double* A22 = A + (long long)j1 * lda + j1;
        const int nb1 = min(CHOL_NB, rem);           // width of the next block column
        if (lookahead && rem > nb1) {
            // While the trailing update is the bottleneck (large remainders) two block columns are applied together: a rank-128
            // update moves 320 KB through L2 per 128 x 64 tile for 2.1 MFLOP and runs at ~55 % of the DMMA peak, the rank-256
            // update of a pair moves 512 KB for 4.2 MFLOP.  Even step of a pair: only the next block column is brought up to
            // date (look-ahead, K = 128) and the bulk is deferred; odd step: look-ahead and bulk with K = 256 over both panels,
            // which are adjacent columns of A.  Small remainders (bulk hidden behind the critical path) keep K = 128.
            // The look-ahead of the odd step covers TWO block columns, so that the even step of the next pair (which brings the
            // second of them up to date) never touches what the paired bulk update is writing and does not have to wait for it.
            // Measured at n = 5000: 4.03 ms paired against 4.04 ms unpaired -- the rank-256 CTAs live twice as long, and the
            // one-CTA diagonal-block kernel (215 KB of shared memory: it needs a whole SM) then waits ~25 us for an SM to drain
            // at every other step, which returns what the better GEMM rate gains.  Opt-in: tuning key 6 = 1.
            const bool pair_second = pair_open;
            const bool pair_first = !pair_open && g_tuning[6] == 1 && jb == CHOL_NB && rem > PAIR_MIN_REM;
            const int la_cols = pair_second ? min(2 * CHOL_NB, rem) : nb1;       // width of the look-ahead
            const int kcols = pair_second ? 2 * CHOL_NB : jb;                    // K of this step's updates
            const double* P = pair_second ? A21 - CHOL_NB : A21;                 // [rem x kcols] panel(s), row stride lda
            // (a) next block column(s) on the main stream: A22[:, 0:la_cols] -= P . P[0:la_cols]^T
            if (!pair_first) PPBO_CUDA_CHECK(cudaEventRecord(g_chol.panel_done[b & 1], st));     // panels final: (b) may start
            if (side_busy && !pair_first) PPBO_CUDA_CHECK(cudaStreamWaitEvent(st, g_chol.rest_done[last_rest], 0));
            rc = (la_cols == CHOL_NB && kcols == CHOL_NB && !pdl)
                     ? launch_rowpanel128(P, lda, P, lda, A22, lda, A22, lda, rem, -1.0, 1.0, st) : 0;
            if (rc < 0) return rc;
            if (rc == 0) {
                GemmOperands g{P, lda, 0, P, lda, 0, rem, la_cols, kcols};
                StoreEpilogue ep{A22, lda, 0, -1.0, 1.0, 0, 0};
                g_launch_pdl = pdl;
                rc = launch_gemm_nt(g, ep, 1, st);
                g_launch_pdl = false;
                if (rc) return rc;
            }
            if (tl) { cudaEventRecord(ev_l[b], st); has_l[b] = true; }
            if (pair_first) {
                pair_open = true;                    // bulk deferred to the next step
            } else {
                pair_open = false;
                // (b) the rest of the trailing matrix on the side stream (lower tiles only); it needs the panel(s), not (a)
                PPBO_CUDA_CHECK(cudaStreamWaitEvent(g_chol.side, g_chol.panel_done[b & 1], 0));
                if (tl) { cudaEventRecord(ev_r0[b], g_chol.side); has_r[b] = true; }
                {
                    const int r2 = rem - la_cols;
                    const double* P2 = P + (long long)la_cols * lda;
                    double* C2 = A22 + (long long)la_cols * lda + la_cols;
                    rc = (kcols == CHOL_NB) ? launch_syrk128(P2, lda, C2, lda, r2, g_chol.side) : 0;
                    if (rc < 0) return rc;
                    if (rc == 0) {
                        GemmOperands g{P2, lda, 0, P2, lda, 0, r2, r2, kcols};
                        StoreEpilogue ep{C2, lda, 0, -1.0, 1.0, 1, 0};
                        rc = launch_gemm_nt(g, ep, 1, g_chol.side);
                        if (rc) return rc;
                    }
                }
                last_rest ^= 1;
                PPBO_CUDA_CHECK(cudaEventRecord(g_chol.rest_done[last_rest], g_chol.side));
                if (tl) cudaEventRecord(ev_r1[b], g_chol.side);
                side_busy = true;
            }
        } else {
            if (side_busy) {
                PPBO_CUDA_CHECK(cudaStreamWaitEvent(st, g_chol.rest_done[last_rest], 0));
                side_busy = false;
            }
            GemmOperands g{A21, lda, 0, A21, lda, 0, rem, rem, jb};
            StoreEpilogue ep{A22, lda, 0, -1.0, 1.0, 1, 0};
            g_launch_pdl = pdl;
            rc = launch_gemm_nt(g, ep, 1, st);
            g_launch_pdl = false;
            if (rc) return rc;
        }
    }
    if (side_busy) PPBO_CUDA_CHECK(cudaStreamWaitEvent(st, g_chol.rest_done[last_rest], 0));
    if (tl) {
        cudaEventRecord(ev_end, st);
        cudaDeviceSynchronize();
        float t_end = 0;
        cudaEventElapsedTime(&t_end, ev_d[0], ev_end);
        if (nblk > 20) {
            long long h[32];
            cudaMemcpy(h, g_potf2_dbg, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[potf2 phases, block 20, SM clocks] load %lld | ", h[1] - h[0]);
            for (int sb = 0; sb < 4; ++sb) fprintf(stderr, "sub%d: warp potrf %lld, inverse || solve+trailing %lld | ", sb, h[2 + 2 * sb] - h[1 + 2 * sb], h[3 + 2 * sb] - h[2 + 2 * sb]);
            fprintf(stderr, "tail: last block row of the inverse %lld | store %lld | total %lld\n", h[11] - h[10], h[13] - h[12], h[13] - h[0]);
            for (int sb = 0; sb < 4; ++sb)
                fprintf(stderr, "[warp potrf32 sub%d] enter +%lld | load %lld | columns %lld | scale+store %lld | to barrier exit %lld\n", sb,
                        h[16 + 4 * sb] - h[1 + 2 * sb], h[17 + 4 * sb] - h[16 + 4 * sb], h[18 + 4 * sb] - h[17 + 4 * sb],
                        h[19 + 4 * sb] - h[18 + 4 * sb], h[2 + 2 * sb] - h[19 + 4 * sb]);
        }
        fprintf(stderr, "[potrf timeline n=%d] total %.1f us; per step: start | potf2 | panel+lookahead-issue | rest start..end (us)\n", n, t_end * 1e3);
        for (int b = 0; b < nblk; ++b) {
            float t0 = 0, t1 = 0, t2 = 0, r0 = 0, r1 = 0;
            cudaEventElapsedTime(&t0, ev_d[0], ev_d[b]);
            cudaEventElapsedTime(&t1, ev_d[b], ev_p[b]);
            if (has_l[b]) cudaEventElapsedTime(&t2, ev_p[b], ev_l[b]);
            if (has_r[b]) { cudaEventElapsedTime(&r0, ev_d[0], ev_r0[b]); cudaEventElapsedTime(&r1, ev_d[0], ev_r1[b]); }
            fprintf(stderr, "  b=%2d  %7.1f | %5.1f | %5.1f | %7.1f .. %7.1f (%.1f)\n", b, t0 * 1e3, t1 * 1e3, t2 * 1e3, r0 * 1e3, r1 * 1e3, (r1 - r0) * 1e3);
        }
    }
    if (st != caller) {
        PPBO_CUDA_CHECK(cudaEventRecord(g_chol.join, st));
        PPBO_CUDA_CHECK(cudaStreamWaitEvent(caller, g_chol.join, 0));
    }
    return PPBO_OK;
}

// ------------------------------------------------------------------------------------------- TRSV (one RHS)
// forward step for block J: every CTA recomputes y_J = inv(L_JJ) t_J (cheap), CTA 0 stores it, then each warp
// updates rows below:  t[i] -= L[i, J] . y_J
__global__ void __launch_bounds__(256) trsv_fwd_step(const double* __restrict__ L, long long ldl, int n, int j0, int jb,
                                                     const double* __restrict__ dinv_b, double* __restrict__ t) {
    __shared__ double tj[CHOL_NB], yj[CHOL_NB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < CHOL_NB; i += 256) tj[i] = (i < jb) ? t[j0 + i] : 0.0;
    __syncthreads();
    for (int r = warp; r < jb; r += 8) {
        double s = 0.0;
        for (int k = lane; k <= r; k += 32) s = fma(dinv_b[r * CHOL_NB + k], tj[k], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) yj[r] = s;
    }
    __syncthreads();
    const int j1 = j0 + jb;
    const int rows = n - j1;
    const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    for (int r = gw; r < rows; r += nw) {
        const double* lrow = L + (long long)(j1 + r) * ldl + j0;
        double s = 0.0;
        for (int k = lane; k < jb; k += 32) s = fma(lrow[k], yj[k], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) t[j1 + r] -= s;
    }
    // y_J overwrites t_J only after every CTA has read t_J: done by a second tiny launch (trsv_store) to stay race-free
    if (blockIdx.x == 0) {
        double* ybuf = t + n;   // scratch tail [n, n + NB)
        for (int i = tid; i < jb; i += 256) ybuf[i] = yj[i];
    }
}
__global__ void trsv_commit(double* __restrict__ t, int n, int j0, int jb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < jb) t[j0 + i] = t[n + i];
}
// backward step for block J: x_J = inv(L_JJ)^T y_J ; y[k] -= sum_r L[j0+r][k] x_J[r] for k < j0
__global__ void __launch_bounds__(256) trsv_bwd_step(const double* __restrict__ L, long long ldl, int n, int j0, int jb,
                                                     const double* __restrict__ dinv_b, double* __restrict__ y) {
    __shared__ double yj[CHOL_NB], xj[CHOL_NB];
    const int tid = threadIdx.x;
    for (int i = tid; i < CHOL_NB; i += 256) yj[i] = (i < jb) ? y[j0 + i] : 0.0;
    __syncthreads();
    for (int c = tid; c < jb; c += 256) {          // x[c] = sum_{r >= c} dinv[r][c] y[r]  (coalesced over c)
        double s = 0.0;
        for (int r = c; r < jb; ++r) s = fma(dinv_b[r * CHOL_NB + c], yj[r], s);
        xj[c] = s;
    }
    __syncthreads();
    for (int k = blockIdx.x * 256 + tid; k < j0; k += gridDim.x * 256) {
        double s = 0.0;
#pragma unroll 4
        for (int r = 0; r < jb; ++r) s = fma(L[(long long)(j0 + r) * ldl + k], xj[r], s);
        y[k] -= s;
    }
    if (blockIdx.x == 0) {
        double* xbuf = y + n;
        for (int i = tid; i < jb; i += 256) xbuf[i] = xj[i];
    }
}

// ---- chained single-launch solves ------------------------------------------------------------------------------------
// One CTA per 128-row block; CTA c accumulates its dependencies on earlier blocks as they are published through global
// flags (producer: write, __threadfence, flag; consumer: spin on the flag, then __ldcg), so the critical path of the whole
// solve is nblk x (one 128 x 128 GEMV + publish) instead of nblk kernel launches of latency-bound work.  All CTAs are
// co-resident (nblk <= 128 <= SM count), and every CTA only waits on CTAs with a smaller blockIdx.
constexpr int TRSV_THREADS = 512;

// every lane polls (one transaction per warp); acquire so that the dependency's values are visible afterwards
__device__ __forceinline__ void trsv_wait(const int* flag) {
    int v;
    do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == 0) __nanosleep(20);
    } while (v == 0);
}
__device__ __forceinline__ void trsv_publish(int* flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}

// forward: L y = t (in place).  CTA c: v = t_c - sum_{J<c} L[c,J] y_J ; y_c = inv(L_cc) v.
// Warp w owns rows w, w+16, ... of the block; the L tile of dependency J is requested BEFORE the flag of J is polled, so on
// the critical path (J = c-1) only the 128 values of y_J, one FMA sweep and the 128 x 128 triangular product remain.
// No block-wide barrier inside the dependency loop.
__global__ void __launch_bounds__(TRSV_THREADS) trsv_fwd_chain(const double* __restrict__ L, long long ldl, int n,
                                                               const double* __restrict__ dinv, double* t, int* flags) {
    extern __shared__ double sm[];
    double* dtile = sm;                       // inv(L_cc), 128 x 128 row-major
    double* v = sm + CHOL_NB * CHOL_NB;       // [128]
    const int c = blockIdx.x, j0 = c * CHOL_NB, jb = min(CHOL_NB, n - j0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double* dsrc = dinv + (long long)c * CHOL_NB * CHOL_NB;
    for (int e = tid; e < CHOL_NB * CHOL_NB; e += TRSV_THREADS) dtile[e] = dsrc[e];
    double s[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) s[rr] = 0.0;
    for (int J = 0; J < c; ++J) {
        double l[8][4];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            const int row = min(warp + 16 * rr, jb - 1);                     // clamp: rows >= jb are discarded below
            const double* lr = L + (long long)(j0 + row) * ldl + (long long)J * CHOL_NB + lane;
#pragma unroll
            for (int q = 0; q < 4; ++q) l[rr][q] = lr[32 * q];
        }
        trsv_wait(flags + J);
        const double* yJ = t + (long long)J * CHOL_NB + lane;
        const double y0 = __ldcg(yJ), y1 = __ldcg(yJ + 32), y2 = __ldcg(yJ + 64), y3 = __ldcg(yJ + 96);
#pragma unroll
        for (int rr = 0; rr < 8; ++rr)
            s[rr] = fma(l[rr][0], y0, fma(l[rr][1], y1, fma(l[rr][2], y2, fma(l[rr][3], y3, s[rr]))));
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        double a = s[rr];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        const int row = warp + 16 * rr;
        if (lane == 0) v[row] = (row < jb) ? t[j0 + row] - a : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {          // y_c[row] = sum_{k <= row} dinv[row][k] v[k]
        const int row = warp + 16 * rr;
        double a = 0.0;
        for (int k = lane; k <= row; k += 32) a = fma(dtile[row * CHOL_NB + k], v[k], a);
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && row < jb) t[j0 + row] = a;
    }
    trsv_publish(flags + c);
}

// backward: L^T x = y (in place).  blockIdx b handles block c = nblk-1-b: v = y_c - sum_{K>c} L[K,c]^T x_K ; x_c = inv(L_cc)^T v.
// Thread (g, k): column k of the block, rows g*32 .. g*32+31 of every dependency tile; the 32 x-values a warp needs are one
// coalesced load, broadcast by shuffle.  No block-wide barrier inside the dependency loop.
__global__ void __launch_bounds__(TRSV_THREADS) trsv_bwd_chain(const double* __restrict__ L, long long ldl, int n, int nblk,
                                                               const double* __restrict__ dinv, double* t, int* flags) {
    extern __shared__ double sm[];
    double* dtile = sm;
    double* xv = sm + CHOL_NB * CHOL_NB;      // [128]
    double* part = xv + CHOL_NB;              // [4][128]
    const int c = nblk - 1 - (int)blockIdx.x, j0 = c * CHOL_NB;
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 7, k = tid & 127;     // 4 row groups x 128 columns
    const double* dsrc = dinv + (long long)c * CHOL_NB * CHOL_NB;
    for (int e = tid; e < CHOL_NB * CHOL_NB; e += TRSV_THREADS) dtile[e] = dsrc[e];
    double acc0 = 0.0, acc1 = 0.0;
    const int col = min(j0 + k, n - 1);       // clamp: columns >= n are discarded below
    for (int K = nblk - 1; K > c; --K) {
        const int kb = min(CHOL_NB, n - K * CHOL_NB);
        double l[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int row = min(K * CHOL_NB + g * 32 + r, n - 1);
            l[r] = L[(long long)row * ldl + col];
        }
        trsv_wait(flags + K);
        const int xr = g * 32 + lane;
        const double xmine = (xr < kb) ? __ldcg(t + (long long)K * CHOL_NB + xr) : 0.0;    // rows >= kb contribute nothing
#pragma unroll
        for (int r = 0; r < 32; r += 2) {
            acc0 = fma(l[r], __shfl_sync(0xffffffffu, xmine, r), acc0);
            acc1 = fma(l[r + 1], __shfl_sync(0xffffffffu, xmine, r + 1), acc1);
        }
    }
    part[g * CHOL_NB + k] = acc0 + acc1;
    __syncthreads();
    if (tid < CHOL_NB) {
        const double a = part[tid] + part[CHOL_NB + tid] + part[2 * CHOL_NB + tid] + part[3 * CHOL_NB + tid];
        xv[tid] = (j0 + tid < n) ? t[j0 + tid] - a : 0.0;          // v
    }
    __syncthreads();
    {   // x_c[k] = sum_{r >= k} dinv[r][k] v[r]; thread (g, k) takes rows r = g, g+4, ...
        double a = 0.0;
        for (int r = g; r < CHOL_NB; r += 4)
            if (r >= k) a = fma(dtile[r * CHOL_NB + k], xv[r], a);
        part[g * CHOL_NB + k] = a;
    }
    __syncthreads();
    if (tid < CHOL_NB && j0 + tid < n)
        t[j0 + tid] = part[tid] + part[CHOL_NB + tid] + part[2 * CHOL_NB + tid] + part[3 * CHOL_NB + tid];
    trsv_publish(flags + c);
}

constexpr int TRSV_CHAIN_SMEM = (CHOL_NB * CHOL_NB + 6 * CHOL_NB) * (int)sizeof(double);

// solves (L L^T) x = t in place; t must have room for n + CHOL_NB doubles (scratch tail: flags of the chained kernels)
int potrs_vec(const double* L, long long ldl, int n, const double* dinv, double* t, cudaStream_t st) {
    const int nblk = ceil_div(n, CHOL_NB);
    if (nblk <= 128) {
        static std::once_flag once;
        static cudaError_t err = cudaSuccess;
        std::call_once(once, [] {
            err = cudaFuncSetAttribute(trsv_fwd_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_CHAIN_SMEM);
            if (err == cudaSuccess)
                err = cudaFuncSetAttribute(trsv_bwd_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSV_CHAIN_SMEM);
        });
        PPBO_CUDA_CHECK(err);
        int* flags = reinterpret_cast<int*>(t + n);          // 2 * nblk ints <= 1 KB
        PPBO_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int) * 2 * nblk, st));
        PPBO_CL trsv_fwd_chain<<<nblk, TRSV_THREADS, TRSV_CHAIN_SMEM, st>>>(L, ldl, n, dinv, t, flags);
        PPBO_CL trsv_bwd_chain<<<nblk, TRSV_THREADS, TRSV_CHAIN_SMEM, st>>>(L, ldl, n, nblk, dinv, t, flags + nblk);
        PPBO_LAUNCH_CHECK();
        return PPBO_OK;
    }
    for (int b = 0; b < nblk; ++b) {
        const int j0 = b * CHOL_NB, jb = min(CHOL_NB, n - j0);
        const int rows = n - j0 - jb;
        const int blocks = max(1, min(ceil_div(rows, 8 * 4), PPBO_SM_COUNT * 2));
        PPBO_CL trsv_fwd_step<<<blocks, 256, 0, st>>>(L, ldl, n, j0, jb, dinv + (long long)b * CHOL_NB * CHOL_NB, t);
        PPBO_CL trsv_commit<<<1, CHOL_NB, 0, st>>>(t, n, j0, jb);
    }
    for (int b = nblk - 1; b >= 0; --b) {
        const int j0 = b * CHOL_NB, jb = min(CHOL_NB, n - j0);
        const int blocks = max(1, min(ceil_div(j0, 256), PPBO_SM_COUNT * 2));
        PPBO_CL trsv_bwd_step<<<blocks, 256, 0, st>>>(L, ldl, n, j0, jb, dinv + (long long)b * CHOL_NB * CHOL_NB, t);
        PPBO_CL trsv_commit<<<1, CHOL_NB, 0, st>>>(t, n, j0, jb);
    }
    PPBO_LAUNCH_CHECK();
    return PPBO_OK;
}


// ------------------------------------------------------------------------------------------- block-inverse solves
// A triangular solve with one right-hand side reads 8 n^2 / 2 bytes (100 MB at n = 5000: 16 us of HBM time) but the chained
// kernel above pays ~6 us of dependency latency per 128-row block (40 links: 0.25 ms per direction).  When one factor serves
// many solves (the chord steps of the Laplace / weight-space Newton iterations) the diagonal blocks of L are inverted at
// BI = 1024 granularity once -- by doubling the 128 x 128 inverses of the factorisation,
//   inv [[A, 0], [C, B]] = [[A^-1, 0], [-B^-1 C A^-1, B^-1]]    (three batched NT GEMMs per level, both A^-1 and A^-T kept),
// and a solve becomes n / 1024 steps of two bandwidth-bound matrix-vector kernels on the whole GPU.
constexpr int BI = 1024;

long long blockinv_doubles(int n) {
    const long long nbI = ceil_div(n, BI);
    return nbI * (3LL * BI * BI + BI * BI / 4 + BI);
}

// per 1024-block J: Lpad = lower triangle of L's diagonal block (identity on the padding), Binv / BinvT = block diagonal of the
// 128 x 128 inverses (and transposes) produced by potrf_lower
__global__ void __launch_bounds__(256) blockinv_init_kernel(const double* __restrict__ L, long long ldl, int n,
                                                            const double* __restrict__ dinv, double* __restrict__ Binv,
                                                            double* __restrict__ BinvT, double* __restrict__ Lpad) {
    const int J = blockIdx.y;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int i = e / BI, j = e % BI;
    const int gi = J * BI + i, gj = J * BI + j;
    const long long o = (long long)J * BI * BI + e;
    double lp = (i == j) ? 1.0 : 0.0, bi = lp, bt = lp;
    if (gi < n && gj < n) {
        lp = (j <= i) ? L[(long long)gi * ldl + gj] : 0.0;
        bi = bt = 0.0;
        if ((i >> 7) == (j >> 7)) {
            const double* d = dinv + (long long)(gi >> 7) * CHOL_NB * CHOL_NB;
            bi = d[(i & 127) * CHOL_NB + (j & 127)];
            bt = d[(j & 127) * CHOL_NB + (i & 127)];
        }
    } else if (gi < n || gj < n) {
        lp = bi = bt = 0.0;
    }
    Lpad[o] = lp;
    Binv[o] = bi;
    BinvT[o] = bt;
}

// first_block > 0: the 1024-blocks below it are still valid from an earlier build with the same ceil(n / 1024) (a factor that only
// grew at its end, laplace.cu bordered warm start) and are skipped
int blockinv_build(const double* L, long long ldl, int n, const double* dinv, double* W, cudaStream_t st, int first_block) {
    const int nbI = ceil_div(n, BI);
    const long long BB = (long long)BI * BI;
    if (first_block >= nbI) return PPBO_OK;
    const int nbuild = nbI - first_block;
    double* Binv = W + first_block * BB;
    double* BinvT = W + nbI * BB + first_block * BB;
    double* Lpad = W + 2 * nbI * BB + first_block * BB;
    double* Tt = W + 3 * nbI * BB + first_block * (BB / 4);
    L += (long long)first_block * BI * (ldl + 1);
    dinv += (long long)first_block * (BI / CHOL_NB) * CHOL_NB * CHOL_NB;
    n -= first_block * BI;
    PPBO_CL blockinv_init_kernel<<<dim3(BI * BI / 256, nbuild), 256, 0, st>>>(L, ldl, n, dinv, Binv, BinvT, Lpad);
    PPBO_LAUNCH_CHECK();
    for (int s = CHOL_NB; s < BI; s *= 2) {
        const int ppb = BI / (2 * s), nb = nbuild * ppb;
        const long long pair = 2LL * s * (BI + 1), ss = (long long)s * s;
        int rc;
        {   // Tt = (C A^-1)^T = A^-T . C^T
            GemmOperands g{BinvT, BI, pair, Lpad + (long long)s * BI, BI, pair, s, s, s, ppb, BB, BB};
            StoreEpilogue ep{Tt, s, ss, 1.0, 0.0, 0, 0, 0, ppb * ss};
            if ((rc = launch_gemm_nt(g, ep, nb, st))) return rc;
        }
        {   // X = -B^-1 (C A^-1)   -> lower-left block of the inverse
            GemmOperands g{Binv + (long long)s * (BI + 1), BI, pair, Tt, s, ss, s, s, s, ppb, BB, ppb * ss};
            StoreEpilogue ep{Binv + (long long)s * BI, BI, pair, -1.0, 0.0, 0, 0, 0, BB};
            if ((rc = launch_gemm_nt(g, ep, nb, st))) return rc;
        }
        {   // X^T -> upper-right block of the transposed inverse
            GemmOperands g{Tt, s, ss, Binv + (long long)s * (BI + 1), BI, pair, s, s, s, ppb, ppb * ss, BB};
            StoreEpilogue ep{BinvT + s, BI, pair, -1.0, 0.0, 0, 0, 0, BB};
            if ((rc = launch_gemm_nt(g, ep, nb, st))) return rc;
        }
    }
    return PPBO_OK;
}

// out[r] = sum_c Mat[r][c] v[c] over the triangle (lower: c <= r, upper: r <= c < rows) of one BI x BI block.
// 128 threads per row, two rows per CTA; every thread issues its (up to) four 16-byte loads of the row at once, so the whole
// 4 MB triangle is one DRAM round trip deep (the warp-per-row version walked a 1024-element row in 8 dependent rounds: 15 us).
// The other triangle of Mat holds exact zeros; v is only read inside [c_lo, c_hi) (the tail behind the block may hold anything).
constexpr int BTRI_ROWS = 2;
__global__ void __launch_bounds__(128 * BTRI_ROWS) blocktri_gemv_kernel(const double* __restrict__ Mat, const double* __restrict__ v,
                                                                        double* __restrict__ out, int rows, int upper,
                                                                        const double* __restrict__ skip) {
    __shared__ double part[BTRI_ROWS][4];
    if (skip && *skip != 0.0) return;
    const int sub = threadIdx.x >> 7, t = threadIdx.x & 127, lane = threadIdx.x & 31, w = t >> 5;
    const int r = blockIdx.x * BTRI_ROWS + sub;
    const bool vec = (reinterpret_cast<uintptr_t>(Mat) & 15) == 0;          // workspace carving may leave Mat 8-byte aligned
    double s = 0.0;
    if (r < rows) {
        const double* m = Mat + (long long)r * BI;
        const int c_lo = upper ? r : 0, c_hi = upper ? rows : r + 1;           // columns [c_lo, c_hi)
        double2 a[4];
        double x0[4], x1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 2 * t + 256 * q;                                      // even column of this thread's pair
            const bool p0 = c >= c_lo && c < c_hi, p1 = c + 1 >= c_lo && c + 1 < c_hi;
            if (vec) a[q] = (p0 || p1) ? *reinterpret_cast<const double2*>(m + c) : make_double2(0.0, 0.0);
            else a[q] = make_double2(p0 ? m[c] : 0.0, p1 ? m[c + 1] : 0.0);
            x0[q] = p0 ? v[c] : 0.0;
            x1[q] = p1 ? v[c + 1] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) s = fma(a[q].x, x0[q], fma(a[q].y, x1[q], s));
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) part[sub][w] = s;
    __syncthreads();
    if (t == 0 && r < rows) out[r] = (part[sub][0] + part[sub][1]) + (part[sub][2] + part[sub][3]);
}
// forward sweep: t[r] -= sum_{c < cols} L[r][c0 + c] y[c] for r in [r0, n); warp per row
__global__ void __launch_bounds__(256) blockrow_update_kernel(const double* __restrict__ L, long long ldl, int r0, int n, int c0,
                                                              int cols, const double* __restrict__ y, double* __restrict__ t,
                                                              const double* __restrict__ skip) {
    if (skip && *skip != 0.0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = r0 + blockIdx.x * 8 + warp;
    if (r >= n) return;
    const double* l = L + (long long)r * ldl + c0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = lane;
    for (; c + 96 < cols; c += 128) {
        s0 = fma(l[c], y[c], s0);
        s1 = fma(l[c + 32], y[c + 32], s1);
        s2 = fma(l[c + 64], y[c + 64], s2);
        s3 = fma(l[c + 96], y[c + 96], s3);
    }
    for (; c < cols; c += 32) s0 = fma(l[c], y[c], s0);
    double a = (s0 + s1) + (s2 + s3);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) t[r] -= a;
}
// backward sweep: y[c] -= sum_{r0 <= r < r1} L[r][c] x[r - r0] for c < c1; a CTA owns 32 columns, its 32 warps split the rows
// and keep 8 row segments (256 B each) in flight per warp -- the one-warp-per-128-rows version with two loads in flight was
// latency-bound at ~0.9 TB/s
constexpr int BCOL_WARPS = 32;
__global__ void __launch_bounds__(BCOL_WARPS * 32) blockcol_update_kernel(const double* __restrict__ L, long long ldl, int r0, int r1,
                                                                          int c1, const double* __restrict__ x,
                                                                          double* __restrict__ y, const double* __restrict__ skip) {
    __shared__ double part[BCOL_WARPS][33];
    if (skip && *skip != 0.0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + lane;
    double s[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] = 0.0;
    if (c < c1) {
        for (int r = r0 + warp; r < r1; r += 8 * BCOL_WARPS) {
            double l[8], xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r + u * BCOL_WARPS;
                const bool ok = rr < r1;
                l[u] = ok ? __ldcs(L + (long long)rr * ldl + c) : 0.0;
                xv[u] = ok ? x[rr - r0] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] = fma(l[u], xv[u], s[u]);
        }
    }
    part[warp][lane] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    __syncthreads();
    if (warp == 0 && c < c1) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < BCOL_WARPS; ++w) a += part[w][lane];
        y[c] -= a;
    }
}

// scratch vector of the block-inverse solves (behind the inverses; layout fixed by the n the inverses were built for)
double* blockinv_y(double* W, int n) {
    const int nbI = ceil_div(n, BI);
    return W + nbI * (3LL * BI * BI + BI * BI / 4);
}
// forward half: y = L^-1 t  (t is consumed; y = blockinv_y(W, n))
int potrs_fwd_blockinv(const double* L, long long ldl, int n, double* W, double* t, cudaStream_t st, const double* skip) {
    const int nbI = ceil_div(n, BI);
    const long long BB = (long long)BI * BI;
    const double* Binv = W;
    double* y = blockinv_y(W, n);
    for (int J = 0; J < nbI; ++J) {                     // L y = t
        const int j0 = J * BI, rows = min(BI, n - j0), j1 = j0 + rows;
        PPBO_CL blocktri_gemv_kernel<<<ceil_div(rows, BTRI_ROWS), 128 * BTRI_ROWS, 0, st>>>(Binv + J * BB, t + j0, y + j0, rows, 0, skip);
        if (j1 < n)
            PPBO_CL blockrow_update_kernel<<<ceil_div(n - j1, 8), 256, 0, st>>>(L, ldl, j1, n, j0, rows, y + j0, t, skip);
    }
    PPBO_LAUNCH_CHECK();
    return PPBO_OK;
}
// backward half: t = L^-T y  (y is consumed)
int potrs_bwd_blockinv(const double* L, long long ldl, int n, double* W, double* t, cudaStream_t st, const double* skip) {
    const int nbI = ceil_div(n, BI);
    const long long BB = (long long)BI * BI;
    const double* BinvT = W + nbI * BB;
    double* y = blockinv_y(W, n);
    for (int J = nbI - 1; J >= 0; --J) {                // L^T x = y
        const int j0 = J * BI, rows = min(BI, n - j0);
        PPBO_CL blocktri_gemv_kernel<<<ceil_div(rows, BTRI_ROWS), 128 * BTRI_ROWS, 0, st>>>(BinvT + J * BB, y + j0, t + j0, rows, 1, skip);
        if (j0 > 0)
            PPBO_CL blockcol_update_kernel<<<ceil_div(j0, 32), BCOL_WARPS * 32, 0, st>>>(L, ldl, j0, j0 + rows, j0, t + j0, y, skip);
    }
    PPBO_LAUNCH_CHECK();
    return PPBO_OK;
}
// (L L^T) x = t in place with the block inverses W of blockinv_build
int potrs_vec_blockinv(const double* L, long long ldl, int n, double* W, double* t, cudaStream_t st, const double* skip) {
    int rc = potrs_fwd_blockinv(L, ldl, n, W, t, st, skip);
    if (rc) return rc;
    return potrs_bwd_blockinv(L, ldl, n, W, t, st, skip);
}

// R[nrhs x n] = T . L^-T with the 1024-block inverses (T is destroyed): n / 1024 steps of two GEMMs instead of the n / 128 steps of
// trsm_right_lower_t -- the few-row triangular solve of the bordered warm start (laplace.cu) is launch-bound, not flop-bound
int trsm_right_blockinv(const double* L, long long ldl, int n, const double* W, double* T, long long ldt, double* R, long long ldr,
                        int nrhs, cudaStream_t st) {
    const int nbI = ceil_div(n, BI);
    const long long BB = (long long)BI * BI;
    for (int J = 0; J < nbI; ++J) {
        const int j0 = J * BI, rows = min(BI, n - j0), j1 = j0 + rows;
        {
            GemmOperands g{T + j0, ldt, 0, W + J * BB, BI, 0, nrhs, rows, rows};
            StoreEpilogue ep{R + j0, ldr, 0, 1.0, 0.0, 0, 0, 0};
            int rc = launch_gemm_nt(g, ep, 1, st);
            if (rc) return rc;
        }
        if (j1 < n) {
            GemmOperands g{R + j0, ldr, 0, L + (long long)j1 * ldl + j0, ldl, 0, nrhs, n - j1, rows};
            StoreEpilogue ep{T + j1, ldt, 0, -1.0, 1.0, 0, 0, 0};
            int rc = launch_gemm_nt(g, ep, 1, st);
            if (rc) return rc;
        }
    }
    return PPBO_OK;
}

// ------------------------------------------------------------------------------------------- TRSM (many RHS, one per row)
// X[nrhs x n] <- X . L^-T : column blocks ascending; Y_J = X_J . inv(L_JJ)^T ; X[:, J+1:] -= Y_J . L[J+1:, J]^T
int trsm_right_lower_t(const double* L, long long ldl, int n, const double* dinv, double* X, long long ldx, int nrhs,
                       cudaStream_t st) {
    const int nblk = ceil_div(n, CHOL_NB);
    for (int b = 0; b < nblk; ++b) {
        const int j0 = b * CHOL_NB, jb = min(CHOL_NB, n - j0), j1 = j0 + jb;
        {
            GemmOperands g{X + j0, ldx, 0, dinv + (long long)b * CHOL_NB * CHOL_NB, CHOL_NB, 0, nrhs, jb, jb};
            StoreEpilogue ep{X + j0, ldx, 0, 1.0, 0.0, 0, 0, 1};
            int rc = launch_gemm_nt(g, ep, 1, st);   // in place (CfgWide)
            if (rc) return rc;
        }
        if (j1 < n) {
            GemmOperands g{X + j0, ldx, 0, L + (long long)j1 * ldl + j0, ldl, 0, nrhs, n - j1, jb};
            StoreEpilogue ep{X + j1, ldx, 0, -1.0, 1.0, 0, 0};
            int rc = launch_gemm_nt(g, ep, 1, st);
            if (rc) return rc;
        }
    }
    return PPBO_OK;
}

// ------------------------------------------------------------------------------------------- LU with partial pivoting (log-determinant)
// GPModel.evidence (src/gp_model.py:301-308) takes the log-determinant of the NON-symmetric, possibly indefinite matrix
// I + Sigma Lambda_MAP through scipy.linalg.lu (LAPACK dgetrf: row pivoting by largest magnitude, first index on ties).  Blocked
// right-looking LU with the same pivoting rule: per 64-column panel one CTA factors the panel column by column (the panel stays in
// L2), the row interchanges are applied to the rest of the matrix, U12 = L11^-1 A12 by one thread per column (written transposed as
// well: the K-contiguous operand of the update), and A22 -= L21 U12 on the FP64 tensor pipe (gemm_nt_store_kernel).  Off the hot
// path (hyper-parameter search, off by default: ppbo_numerical_main.py:188-190): written for clarity, ~2/3 n^3 flop.
constexpr int LU_NB = 64;

// one CTA: factor the panel A[j0:n, j0:j0+jb] in place with row pivoting; piv[c] = row chosen for column j0 + c (global index);
// stat[0] += number of interchanges, stat[1] = 1 if an exactly zero pivot was met
__global__ void __launch_bounds__(1024) lu_panel_kernel(double* __restrict__ A, long long lda, int n, int j0, int jb,
                                                        int* __restrict__ piv, int* __restrict__ stat) {
    __shared__ double vmax[32];
    __shared__ int imax[32];
    __shared__ double prow[LU_NB];
    __shared__ int psel;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = 0; c < jb; ++c) {
        const int col = j0 + c;
        // (a) pivot search over rows col .. n-1 : largest |value|, smallest row index on ties
        double best = -1.0;
        int bi = col;
        for (int r = col + tid; r < n; r += 1024) {
            const double v = fabs(A[(long long)r * lda + col]);
            if (v > best) { best = v; bi = r; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { vmax[warp] = best; imax[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 32; ++w)
                if (vmax[w] > best || (vmax[w] == best && imax[w] < bi)) { best = vmax[w]; bi = imax[w]; }
            psel = bi;
            piv[c] = bi;
            if (bi != col) atomicAdd(stat, 1);
            if (best == 0.0) stat[1] = 1;
        }
        __syncthreads();
        const int pr = psel;
        // (b) interchange inside the panel, keep the pivot row in shared memory
        if (tid < jb) {
            double* a = A + (long long)col * lda + j0 + tid;
            double* b = A + (long long)pr * lda + j0 + tid;
            const double va = *a, vb = *b;
            *a = vb;
            *b = va;
            prow[tid] = vb;
        }
        __syncthreads();
        const double inv = 1.0 / prow[c];
        // (c) multipliers and the rank-1 update of the remaining panel columns: thread = (row, 8 columns)
        const int rem = jb - c - 1;
        for (int r = col + 1 + warp; r < n; r += 32) {
            double* row = A + (long long)r * lda + col;
            const double l = row[0] * inv;
            __syncwarp();
            if (lane == 0) row[0] = l;
            for (int k = 1 + lane; k <= rem; k += 32) row[k] = fma(-l, prow[c + k], row[k]);
        }
        __syncthreads();
    }
}
// apply the panel's interchanges to the columns outside the panel (in order)
__global__ void __launch_bounds__(256) lu_swap_kernel(double* __restrict__ A, long long lda, int n, int j0, int jb,
                                                      const int* __restrict__ piv) {
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= j0) col += jb;
    if (col >= n) return;
    for (int c = 0; c < jb; ++c) {
        const int p = piv[c];
        if (p != j0 + c) {
            const double a = A[(long long)(j0 + c) * lda + col], b = A[(long long)p * lda + col];
            A[(long long)(j0 + c) * lda + col] = b;
            A[(long long)p * lda + col] = a;
        }
    }
}
// U12 = L11^-1 A12 (unit lower L11, jb x jb) for the columns right of the panel; Ut[col - j1][k] = U12[k][col]
__global__ void __launch_bounds__(128) lu_u12_kernel(double* __restrict__ A, long long lda, int n, int j0, int jb,
                                                     double* __restrict__ Ut) {
    __shared__ double L11[LU_NB][LU_NB + 1];
    for (int e = threadIdx.x; e < jb * jb; e += blockDim.x) L11[e / jb][e % jb] = A[(long long)(j0 + e / jb) * lda + j0 + e % jb];
    __syncthreads();
    const int j1 = j0 + jb;
    const int col = j1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n) return;
    double u[LU_NB];
#pragma unroll 8
    for (int k = 0; k < LU_NB; ++k) u[k] = (k < jb) ? A[(long long)(j0 + k) * lda + col] : 0.0;
#pragma unroll 4
    for (int k = 0; k < LU_NB; ++k) {
        if (k >= jb) break;
        const double x = u[k];
        for (int r = k + 1; r < jb; ++r) u[r] = fma(-L11[r][k], x, u[r]);
    }
    for (int k = 0; k < jb; ++k) {
        A[(long long)(j0 + k) * lda + col] = u[k];
        Ut[(long long)(col - j1) * LU_NB + k] = u[k];
    }
}
// out[0] = sign of prod_i A[i][i], out[1] = sum_i log |A[i][i]|   (single CTA, fixed order)
__global__ void __launch_bounds__(1024) lu_diag_logdet_kernel(const double* __restrict__ A, long long lda, int n, double* __restrict__ out) {
    __shared__ double red[33];
    double s = 0.0, neg = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const double d = A[(long long)i * lda + i];
        s += log(fabs(d));
        if (d < 0.0) neg += 1.0;
    }
    s = block_sum(s, red);
    neg = block_sum(neg, red);
    if (threadIdx.x == 0) {
        out[0] = (fmod(neg, 2.0) == 0.0) ? 1.0 : -1.0;
        out[1] = s;
    }
}

long long lu_workspace_doubles(int n) { return (long long)n * LU_NB + 8 + LU_NB; }

// in-place LU of A[n x n]; res_h[0] = sign(det U), res_h[1] = log|det U| = log|det A|, res_h[2] = sign of the row permutation
int lu_logdet(double* A, long long lda, int n, double* ws, double* res_h, cudaStream_t st) {
    double* Ut = ws;
    double* outd = ws + (long long)n * LU_NB;
    int* piv = reinterpret_cast<int*>(outd + 4);
    int* stat = piv + LU_NB;
    PPBO_CUDA_CHECK(cudaMemsetAsync(stat, 0, 2 * sizeof(int), st));
    for (int j0 = 0; j0 < n; j0 += LU_NB) {
        const int jb = min(LU_NB, n - j0), j1 = j0 + jb;
        PPBO_CL lu_panel_kernel<<<1, 1024, 0, st>>>(A, lda, n, j0, jb, piv, stat);
        if (n - jb > 0) PPBO_CL lu_swap_kernel<<<ceil_div(n - jb, 256), 256, 0, st>>>(A, lda, n, j0, jb, piv);
        if (j1 < n) {
            PPBO_CL lu_u12_kernel<<<ceil_div(n - j1, 128), 128, 0, st>>>(A, lda, n, j0, jb, Ut);
            PPBO_LAUNCH_CHECK();
            GemmOperands g{A + (long long)j1 * lda + j0, lda, 0, Ut, LU_NB, 0, n - j1, n - j1, jb};
            StoreEpilogue ep{A + (long long)j1 * lda + j1, lda, 0, -1.0, 1.0, 0, 0, 0};
            int rc = launch_gemm_nt(g, ep, 1, st);
            if (rc) return rc;
        }
    }
    PPBO_CL lu_diag_logdet_kernel<<<1, 1024, 0, st>>>(A, lda, n, outd);
    PPBO_LAUNCH_CHECK();
    double h[2];
    int sh[2];
    PPBO_CUDA_CHECK(readback().add(h, outd, sizeof(h), st));
    PPBO_CUDA_CHECK(readback().add(sh, stat, sizeof(sh), st));
    PPBO_CUDA_CHECK(readback().finish(st));
    res_h[0] = h[0];
    res_h[1] = h[1];
    res_h[2] = (sh[0] % 2 == 0) ? 1.0 : -1.0;
    return sh[1] ? 1 : PPBO_OK;          // 1: exactly singular
}

}  // namespace ppbo

// =========================================================================================== C ABI
using namespace ppbo;

extern "C" int ppbo_version(void) { return 100; }
extern "C" long long ppbo_launch_count(void) { return g_launch_count; }
extern "C" int ppbo_set_tuning(int key, int value) {
    PPBO_REQUIRE(key >= 0 && key < 16, "tuning key");
    g_tuning[key] = value;
    return PPBO_OK;
}
extern "C" const char* ppbo_last_error(void) { return g_err; }
/* Marks the calling host thread as background work: the streams the library creates for it (Cholesky critical path / bulk) get the
 * lowest priority.  Must be called before the thread's first factorisation. */
extern "C" int ppbo_set_thread_background(int on) {
    g_thread_background = on != 0;
    return PPBO_OK;
}
extern "C" int ppbo_device_sm_count(int dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return v;
}

extern "C" int ppbo_gemm_nt(const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
                            int M, int N, int K, double alpha, double beta, void* stream) {
    PPBO_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
    GemmOperands g{A, lda, 0, B, ldb, 0, M, N, K};
    StoreEpilogue ep{C, ldc, 0, alpha, beta, 0, 0};
    return launch_gemm_nt(g, ep, 1, (cudaStream_t)stream);
}

/* debug / tuning: same as ppbo_gemm_nt with an explicit tile configuration (0 big, 1 small, 2 tall3, 3 sq16, 4 sq3, 5 tall4) */
extern "C" int ppbo_gemm_nt_cfg(int cfg, const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc,
                                int M, int N, int K, double alpha, double beta, void* stream) {
    GemmOperands g{A, lda, 0, B, ldb, 0, M, N, K};
    StoreEpilogue ep{C, ldc, 0, alpha, beta, 0, 0, 0};
    ep.vec_ok = aligned16(C) && ldc % 2 == 0;
    PPBO_REQUIRE(operands_vec2(g), "tuning entry needs 16-byte aligned operands");
    cudaStream_t st = (cudaStream_t)stream;
    switch (cfg) {
        case 0: return launch_store_cfg<CfgBig>(g, ep, 1, st);
        case 1: return launch_store_cfg<CfgSmall>(g, ep, 1, st);
        case 2: return launch_store_cfg<CfgTall3>(g, ep, 1, st);
        case 3: return launch_store_cfg<CfgSq16>(g, ep, 1, st);
        case 4: return launch_store_cfg<CfgSq3>(g, ep, 1, st);
        case 5: return launch_store_cfg<CfgTall4>(g, ep, 1, st);
    }
    PPBO_REQUIRE(false, "unknown configuration");
}

extern "C" long long ppbo_potrf_workspace_bytes(int n) { return potrf_dinv_doubles(n) * 8 + 16; }

extern "C" int ppbo_potrf_lower(double* A, long long lda, int n, void* workspace, long long workspace_bytes,
                                int* info_h, void* stream) {
    PPBO_REQUIRE(n >= 0 && lda >= n, "bad matrix shape");
    PPBO_REQUIRE(workspace_bytes >= ppbo_potrf_workspace_bytes(n), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* dinv = (double*)workspace;
    int* info_d = (int*)(dinv + potrf_dinv_doubles(n));
    int rc = potrf_lower(A, lda, n, dinv, info_d, st);
    if (rc) return rc;
    int info = 0;
    PPBO_CUDA_CHECK(readback().add(&info, info_d, sizeof(int), st));
    PPBO_CUDA_CHECK(readback().finish(st));
    if (info_h) *info_h = info;
    return info;
}

extern "C" int ppbo_trsm_right_lower(const double* L, long long ldl, int n, double* X, long long ldx, int nrhs,
                                     int trans, void* workspace, long long workspace_bytes, void* stream) {
    PPBO_REQUIRE(trans == 0, "only X <- X L^-T is provided (every use on the PPBO path has this form)");
    PPBO_REQUIRE(workspace_bytes >= ppbo_potrf_workspace_bytes(n), "workspace must be the one ppbo_potrf_lower filled");
    return trsm_right_lower_t(L, ldl, n, (const double*)workspace, X, ldx, nrhs, (cudaStream_t)stream);
}

extern "C" int ppbo_potrs_vec(const double* L, long long ldl, int n, double* x, void* workspace, long long workspace_bytes,
                              void* stream) {
    PPBO_REQUIRE(workspace_bytes >= ppbo_potrf_workspace_bytes(n), "workspace must be the one ppbo_potrf_lower filled");
    return potrs_vec(L, ldl, n, (const double*)workspace, x, (cudaStream_t)stream);
}

namespace ppbo { int set_identity(double* A, long long ld, int n, cudaStream_t st); }

/* out = (L L^T)^-1 from the factor ppbo_potrf_lower left in A (lower) and its workspace: Y = I L^-T, out = Y Y^T.
 * work: n*n doubles.  Replaces misc.pd_inverse (src/misc.py:96-100) for the public attributes that are explicit inverses. */
extern "C" long long ppbo_blockinv_bytes(int n) { return n > 0 ? blockinv_doubles(n) * 8 : 0; }

extern "C" int ppbo_blockinv_build(const double* L, long long ldl, int n, const void* potrf_workspace, void* blockinv,
                                   long long blockinv_bytes, void* stream) {
    PPBO_REQUIRE(n >= 1 && ldl >= n, "shape");
    PPBO_REQUIRE(blockinv_bytes >= ppbo_blockinv_bytes(n), "block-inverse workspace too small");
    return blockinv_build(L, ldl, n, reinterpret_cast<const double*>(potrf_workspace), reinterpret_cast<double*>(blockinv),
                          (cudaStream_t)stream);
}

extern "C" int ppbo_potrs_vec_blockinv(const double* L, long long ldl, int n, void* blockinv, long long blockinv_bytes,
                                       double* x, void* stream) {
    PPBO_REQUIRE(n >= 1 && ldl >= n, "shape");
    PPBO_REQUIRE(blockinv_bytes >= ppbo_blockinv_bytes(n), "block-inverse workspace too small");
    return potrs_vec_blockinv(L, ldl, n, reinterpret_cast<double*>(blockinv), x, (cudaStream_t)stream);
}

extern "C" int ppbo_potri_lower(const double* L, long long ldl, int n, void* workspace, long long workspace_bytes, double* work,
                                double* out, long long ldo, void* stream) {
    PPBO_REQUIRE(n >= 0 && ldl >= n && ldo >= n, "shape");
    PPBO_REQUIRE(workspace_bytes >= ppbo_potrf_workspace_bytes(n), "workspace must be the one ppbo_potrf_lower filled");
    if (n == 0) return PPBO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = set_identity(work, n, n, st))) return rc;
    if ((rc = trsm_right_lower_t(L, ldl, n, (const double*)workspace, work, n, n, st))) return rc;     // Y = L^-T (upper)
    GemmOperands g{work, n, 0, work, n, 0, n, n, n};
    StoreEpilogue ep{out, ldo, 0, 1.0, 0.0, 0, 0, 0};
    return launch_gemm_nt(g, ep, 1, st);
}

extern "C" long long ppbo_lu_workspace_bytes(int n) { return n > 0 ? lu_workspace_doubles(n) * 8 : 0; }

/* In-place LU with partial pivoting of A[n x n]; result_h[3] = { sign(det U), log|det A|, sign of the row permutation }.
 * Replaces scipy.linalg.lu + numpy.linalg.slogdet in GPModel.evidence (src/gp_model.py:303-308).  Returns 1 when a pivot is
 * exactly zero. */
extern "C" int ppbo_lu_logdet(double* A, long long lda, int n, void* workspace, long long workspace_bytes, double* result_h,
                              void* stream) {
    PPBO_REQUIRE(n >= 1 && lda >= n && result_h != nullptr, "shape");
    PPBO_REQUIRE(workspace_bytes >= ppbo_lu_workspace_bytes(n), "workspace too small");
    return lu_logdet(A, lda, n, reinterpret_cast<double*>(workspace), result_h, (cudaStream_t)stream);
}

extern "C" int ppbo_gemv(const double* A, long long lda, int M, int N, const double* x, double* y, void* stream) {
    return gemv(A, lda, M, N, x, y, (cudaStream_t)stream);
}
