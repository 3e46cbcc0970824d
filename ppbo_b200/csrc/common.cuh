// ppbo_b200 -- shared definitions for the sm_100a kernels (FP64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define PPBO_MAX_D 64          // max problem dimension carried by value in kernel params
#define PPBO_SM_COUNT 148      // B200

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

}  // namespace ppbo
