// Shared helpers for the factk kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/factk.h"

namespace factk {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// Function attributes (opt-in shared memory) are per device: remember which devices of this process were configured.
inline bool first_use_on_device(unsigned long long& mask) {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    const bool first = (mask & bit) == 0;
    mask |= bit;
    return first;
}

#define FACTK_REQUIRE(cond, ...)                  \
    do {                                          \
        if (!(cond)) {                            \
            ::factk::set_error(__VA_ARGS__);      \
            return FACTK_ERR_ARG;                 \
        }                                         \
    } while (0)

__device__ __forceinline__ float ld_elem(const void* p, int dtype, size_t i) {
    if (dtype == FACTK_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    return reinterpret_cast<const float*>(p)[i];
}

__device__ __forceinline__ void st_elem(void* p, int dtype, size_t i, float v) {
    if (dtype == FACTK_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(p)[i] = v;
}

// 4 consecutive elements starting at element index i (i % 4 == 0 and base 16B-aligned required).
__device__ __forceinline__ float4 ld_vec4(const void* p, int dtype, size_t i) {
    if (dtype == FACTK_BF16) {
        uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
}

__device__ __forceinline__ void st_vec4(void* p, int dtype, size_t i, float4 v) {
    if (dtype == FACTK_BF16) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = u;
    } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide reductions for blockDim.x <= 1024 (sm must hold 32 floats).
__device__ __forceinline__ float block_sum(float v, float* sm) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    float r = (lane < nw) ? sm[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* sm) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    float r = (lane < nw) ? sm[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// Fixed-order sum of per-(video, chunk) partial results (train.cu): out[(b)][i / K][i % K] = alpha * sum (+ out).
void launch_partial_reduce(const float* ws, size_t pstride, int psz, int K, float* out, int ldo, long long out_bstride, int B, int slot,
                           const int32_t* len, int nchunk, int rows_per_chunk, float alpha, int accumulate, cudaStream_t st);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace factk
