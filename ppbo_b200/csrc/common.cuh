// ppbo_b200 -- shared definitions for the sm_100a kernels (FP64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define PPBO_MAX_D 64          // max problem dimension carried by value in kernel params
#define PPBO_SM_COUNT 148      // B200
#define PPBO_MAX_POINTS 64     // points per ppbo_mu_pred_points launch (speculative window of the differential evolution)

#define PPBO_OK 0
#define PPBO_ERR_ARG (-1)
#define PPBO_ERR_CUDA (-2)

namespace ppbo {

void set_error(const char* fmt, ...);
extern long long g_launch_count;   // kernels launched by this library (bench.py reports it as gpu_launches)
#define PPBO_CL ++ppbo::g_launch_count,

#define PPBO_CUDA_CHECK(expr)                                                            \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            ppbo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return PPBO_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

#define PPBO_LAUNCH_CHECK() PPBO_CUDA_CHECK(cudaGetLastError())

#define PPBO_REQUIRE(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            ppbo::set_error("bad argument: %s (%s)", msg, #cond); \
            return PPBO_ERR_ARG;                                  \
        }                                                         \
    } while (0)

// Readback of a few scalars through PINNED host memory.  cudaMemcpyAsync into pageable memory (a stack variable) only returns when
// the stream has reached the copy, and the driver keeps a context lock while it waits: the launches of every other host thread
// stall for that time.  Two host-driven fits on two threads paid for this with each other's batch lengths (weight-space fit: 6 ms
// next to the GP fit against 3 ms alone).  With a pinned destination the call returns at once and the wait happens in
// cudaStreamSynchronize.  One staging buffer per host thread; add() .. add() finish() must not be interleaved with another stream's.
struct PinnedReadback {
    static constexpr size_t CAP = 1 << 16;
    char* buf = nullptr;
    size_t used = 0;
    struct Item { void* dst; size_t off, bytes; } items[16];
    int n = 0;
    cudaError_t add(void* dst, const void* src_dev, size_t bytes, cudaStream_t st) {
        if (!buf && cudaHostAlloc(reinterpret_cast<void**>(&buf), CAP, cudaHostAllocDefault) != cudaSuccess) {
            buf = nullptr;
            (void)cudaGetLastError();
        }
        const size_t need = (bytes + 15) & ~size_t(15);
        if (!buf || n == 16 || used + need > CAP) return cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st);   // pageable fallback
        const cudaError_t e = cudaMemcpyAsync(buf + used, src_dev, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) {                 // (a failed call leaves nothing behind for the next finish())
            items[n++] = Item{dst, used, bytes};
            used += need;
        }
        return e;
    }
    cudaError_t finish(cudaStream_t st) {
        const cudaError_t e = cudaStreamSynchronize(st);
        for (int i = 0; i < n; ++i) memcpy(items[i].dst, buf + items[i].off, items[i].bytes);
        n = 0;
        used = 0;
        return e;
    }
};
PinnedReadback& readback();        // the calling host thread's staging buffer (linalg.cu)

// Stream-ordered scratch.  The device's default memory pool hands its memory back to the driver at every synchronisation unless a
// release threshold is set; every scratch request after a sync then pays a real allocation (~0.5 ms).  The first call per device
// sets the threshold to "keep everything".
cudaError_t malloc_async(void** p, size_t bytes, cudaStream_t st);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// deterministic block-wide sum of one double per thread (blockDim.x multiple of 32, <= 1024)
__device__ inline double block_sum(double v, double* smem33) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) smem33[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (l < nw) ? smem33[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (l == 0) smem33[32] = r;
    }
    __syncthreads();
    return smem33[32];
}

// NV block-wide sums at once (fixed order: shuffle tree inside a warp, warps in index order): v[] holds the totals in EVERY thread on
// return.  smem: 32 * NV doubles.  Three barriers instead of two per value (block_sum): the single-CTA decision kernels of the
// fits spend most of their time in barriers otherwise.  blockDim.x >= NV.
template <int NV>
__device__ inline void block_sum_multi(double (&v)[NV], double* smem) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    __syncthreads();
    if (l == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[w * NV + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < NV) {                     // thread i adds the warps' partial sums of value i (in warp order)
        double r = 0.0;
        for (int k = 0; k < nw; ++k) r += smem[k * NV + threadIdx.x];
        smem[threadIdx.x] = r;                  // row 0 is warp 0's partials: thread i only reads column i of it, before this write
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = smem[i];
}

}  // namespace ppbo
