// Shared device/host helpers for the dpde_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "dpde_b200.h"

namespace dpde {

// ---- host-side error plumbing (no exceptions across the C ABI) -----------------------------------------
char* error_buffer();  // thread-local, defined in update.cu

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(DPDE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return DPDE_OK;
}

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device (update.cu)

constexpr int kThreads = 256;
constexpr int kMaxPartials = 4096;  // upper bound on the reduce grid (per-CTA partial-sum slots)

// ---- device helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_any(const void* p, int dtype, int64_t i) {
    if (dtype == DPDE_F32) return (double)__ldg(reinterpret_cast<const float*>(p) + i);
    if (dtype == DPDE_F64) return __ldg(reinterpret_cast<const double*>(p) + i);
    return (double)__ldg(reinterpret_cast<const unsigned char*>(p) + i);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum three per-thread doubles over a CTA of NT threads; result valid in thread 0.  `scratch` holds 3 * (NT/32) doubles.
template <int NT = kThreads>
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double* scratch) {
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        scratch[warp * 3 + 0] = a;
        scratch[warp * 3 + 1] = b;
        scratch[warp * 3 + 2] = c;
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int nw = NT / 32;
        a = lane < nw ? scratch[lane * 3 + 0] : 0.0;
        b = lane < nw ? scratch[lane * 3 + 1] : 0.0;
        c = lane < nw ? scratch[lane * 3 + 2] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
        c = warp_sum(c);
    }
    __syncthreads();
}

}  // namespace dpde
