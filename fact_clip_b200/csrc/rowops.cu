// Row-local operations: class-softmax splice, LayerNorm, L2 normalise, row softmax, row gather.
// One warp per row; rows are [B][slot][ld] with len[b] valid rows per video.
#include "common.cuh"

namespace factk {

constexpr int ROWS_PER_CTA = 8;   // 8 warps

__device__ __forceinline__ bool row_of_warp(int B, int slot, const int32_t* len, int& b, int& t) {
    const long row = (long)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= (long)B * slot) return false;
    b = (int)(row / slot);
    t = (int)(row % slot);
    return t < (len ? len[b] : slot);
}

// Block.process_feature (models/blocks.py:195-202) + the TDU argmax (blocks.py:420-421).
__global__ void __launch_bounds__(256) softmax_splice_kernel(void* X, int dtype, int B, int slot, const int32_t* len,
                                                             int ld, int H, int C, float* clogit, int32_t* pred) {
    int b, t;
    if (!row_of_warp(B, slot, len, b, t)) return;
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)b * slot + t;
    const size_t base = row * (size_t)ld + (H - C);
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, ld_elem(X, dtype, base + c));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += __expf(ld_elem(X, dtype, base + c) - mx);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float best = -1.f;
    int besti = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float l = ld_elem(X, dtype, base + c);
        const float p = __expf(l - mx) * inv;
        clogit[row * (size_t)C + c] = l;
        st_elem(X, dtype, base + c, p);
        if (p > best) { best = p; besti = c; }   // strict > keeps the first index within a lane
    }
    if (pred) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) pred[row] = besti;
    }
}

// The same for C <= 32 * NV classes with the row's class tail held in registers: one load, one exp and one store per
// element (the generic kernel above re-reads the tail three times and is instruction-issue bound, 90 % issue-active in
// ncu on the frame-level launches).
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, int NV>
__global__ void __launch_bounds__(256) splice_small_kernel(T* __restrict__ X, int B, int slot, const int32_t* __restrict__ len,
                                                           int ld, int H, int C, float* __restrict__ clogit,
                                                           int32_t* __restrict__ pred) {
    int b, t;
    if (!row_of_warp(B, slot, len, b, t)) return;
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)b * slot + t;
    T* x = X + row * (size_t)ld + (H - C);
    float* cl = clogit + row * (size_t)C;
    float v[NV];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < C ? to_f<T>(x[c]) : -INFINITY;
        mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    float e[NV], sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        e[i] = __expf(v[i] - mx);          // exp(-inf) = 0 for the padding lanes
        sum += e[i];
    }
    const float inv = 1.f / warp_sum(sum);
    float best = -1.f;
    int besti = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
            const float p = e[i] * inv;
            cl[c] = v[i];
            x[c] = from_f<T>(p);
            if (p > best) { best = p; besti = c; }   // strict > keeps the first index within a lane
        }
    }
    if (pred) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        if (lane == 0) pred[row] = besti;
    }
}

// splice_small_kernel with RW consecutive rows per warp: all the loads of the RW rows are issued before the first reduction, so a
// warp keeps RW * NV loads in flight instead of NV (the frame-level launches of the one-row kernel ran at a third of the HBM rate:
// 64 warps x 3 short loads per SM do not cover the memory latency).  Same arithmetic, same results.
template <typename T, int NV, int RW>
__global__ void __launch_bounds__(256) splice_multi_kernel(T* __restrict__ X, int B, int slot, const int32_t* __restrict__ len,
                                                           int ld, int H, int C, float* __restrict__ clogit,
                                                           int32_t* __restrict__ pred) {
    const long row0 = ((long)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5)) * RW;
    if (row0 >= (long)B * slot) return;
    const int b = (int)(row0 / slot), t0 = (int)(row0 % slot);          // slot % RW == 0: a group never straddles two videos
    const int nvalid = min(RW, (len ? min(len[b], slot) : slot) - t0);
    if (nvalid <= 0) return;
    const int lane = threadIdx.x & 31;
    T* x0 = X + (size_t)row0 * (size_t)ld + (H - C);
    float v[RW][NV];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            v[r][i] = (r < nvalid && c < C) ? to_f<T>(x0[(size_t)r * ld + c]) : -INFINITY;
        }
    float mx[RW], inv[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        mx[r] = v[r][0];
#pragma unroll
        for (int i = 1; i < NV; ++i) mx[r] = fmaxf(mx[r], v[r][i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < RW; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
    float e[RW][NV];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        inv[r] = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            e[r][i] = __expf(v[r][i] - mx[r]);          // exp(-inf) = 0 for the padding lanes
            inv[r] += e[r][i];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < RW; ++r) inv[r] += __shfl_xor_sync(0xffffffffu, inv[r], o);
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        if (r >= nvalid) break;
        const float iv = 1.f / inv[r];
        const size_t row = (size_t)row0 + r;
        float* cl = clogit + row * (size_t)C;
        float best = -1.f;
        int besti = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float p = e[r][i] * iv;
                cl[c] = v[r][i];
                x0[(size_t)r * ld + c] = from_f<T>(p);
                if (p > best) { best = p; besti = c; }   // strict > keeps the first index within a lane
            }
        }
        if (pred) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            }
            if (lane == 0) pred[row] = besti;
        }
    }
}

__global__ void __launch_bounds__(256) layernorm_kernel(const void* X, int x_dtype, int ldx, const void* R, int r_dtype,
                                                        int ldr, const float* w, const float* bb, float eps, int relu,
                                                        void* Y, int y_dtype, int ldy, int B, int slot,
                                                        const int32_t* len, int E) {
    int b, t;
    if (!row_of_warp(B, slot, len, b, t)) return;
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)b * slot + t;
    float s = 0.f;
    for (int c = lane; c < E; c += 32) {
        float v = ld_elem(X, x_dtype, row * ldx + c);
        if (R) v += ld_elem(R, r_dtype, row * ldr + c);
        s += v;
    }
    const float mu = warp_sum(s) / E;
    float q = 0.f;
    for (int c = lane; c < E; c += 32) {
        float v = ld_elem(X, x_dtype, row * ldx + c);
        if (R) v += ld_elem(R, r_dtype, row * ldr + c);
        q += (v - mu) * (v - mu);
    }
    const float rstd = rsqrtf(warp_sum(q) / E + eps);
    for (int c = lane; c < E; c += 32) {
        float v = ld_elem(X, x_dtype, row * ldx + c);
        if (R) v += ld_elem(R, r_dtype, row * ldr + c);
        v = (v - mu) * rstd * w[c] + bb[c];
        if (relu) v = fmaxf(v, 0.f);
        st_elem(Y, y_dtype, row * ldy + c, v);
    }
}

// LayerNorm with the row in registers (E = 32 * NV columns): one load per element instead of three dtype-dispatched passes.
// Same two-pass mean / variance arithmetic as the generic kernel.
template <typename TX, typename TY, int NV>
__global__ void __launch_bounds__(256) layernorm_small_kernel(const TX* X, int ldx, const float* R, int ldr,   // Y may alias X or R
                                                              const float* __restrict__ w, const float* __restrict__ bb, float eps,
                                                              int relu, TY* Y, int ldy, int B, int slot,
                                                              const int32_t* __restrict__ len) {
    int b, t;
    if (!row_of_warp(B, slot, len, b, t)) return;
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)b * slot + t;
    constexpr int E = 32 * NV;
    const TX* x = X + row * (size_t)ldx;
    float v[NV], s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = to_f<TX>(x[lane + 32 * i]);
        if (R) v[i] += R[row * (size_t)ldr + lane + 32 * i];
        s += v[i];
    }
    const float mu = warp_sum(s) / E;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) q += (v[i] - mu) * (v[i] - mu);
    const float rstd = rsqrtf(warp_sum(q) / E + eps);
    TY* y = Y + row * (size_t)ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        float o = (v[i] - mu) * rstd * w[c] + bb[c];
        if (relu) o = fmaxf(o, 0.f);
        y[c] = from_f<TY>(o);
    }
}

__global__ void __launch_bounds__(256) l2norm_kernel(const void* X, int x_dtype, int ldx, void* Y, int y_dtype, int ldy,
                                                     int B, int slot, const int32_t* len, int E, float eps) {
    int b, t;
    if (!row_of_warp(B, slot, len, b, t)) return;
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)b * slot + t;
    float q = 0.f;
    for (int c = lane; c < E; c += 32) {
        const float v = ld_elem(X, x_dtype, row * ldx + c);
        q += v * v;
    }
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(q)), eps);
    for (int c = lane; c < E; c += 32) st_elem(Y, y_dtype, row * ldy + c, ld_elem(X, x_dtype, row * ldx + c) * inv);
}

__global__ void __launch_bounds__(256) row_softmax_kernel(const float* L, int ldl, float* P, int ldp, int B, int slot,
                                                          const int32_t* len, int M, float scale, __nv_bfloat16* P16,
                                                          int ldp16, int pad16) {
    int b, t;
    if (!row_of_warp(B, slot, len, b, t)) return;
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)b * slot + t;
    float mx = -INFINITY;
    for (int c = lane; c < M; c += 32) mx = fmaxf(mx, L[row * ldl + c] * scale);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < M; c += 32) sum += __expf(L[row * ldl + c] * scale - mx);
    const float inv = 1.f / warp_sum(sum);
    for (int c = lane; c < M; c += 32) {
        const float pv = __expf(L[row * ldl + c] * scale - mx) * inv;
        P[row * ldp + c] = pv;
        if (P16) P16[row * ldp16 + c] = __float2bfloat16_rn(pv);
    }
    if (P16)
        for (int c = M + lane; c < pad16; c += 32) P16[row * ldp16 + c] = __float2bfloat16_rn(0.f);
}

// Row softmax for M <= 32 * NV columns with the row held in registers (same arithmetic as above, one load / exp per element).
template <int NV>
__global__ void __launch_bounds__(256) row_softmax_small_kernel(const float* __restrict__ L, int ldl, float* __restrict__ P, int ldp,
                                                                int B, int slot, const int32_t* __restrict__ len, int M, float scale,
                                                                __nv_bfloat16* __restrict__ P16, int ldp16, int pad16) {
    int b, t;
    if (!row_of_warp(B, slot, len, b, t)) return;
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)b * slot + t;
    const float* l = L + row * (size_t)ldl;
    float v[NV], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < M ? l[c] * scale : -INFINITY;
        mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) { v[i] = __expf(v[i] - mx); sum += v[i]; }
    const float inv = 1.f / warp_sum(sum);
    float* p = P + row * (size_t)ldp;
    __nv_bfloat16* p16 = P16 ? P16 + row * (size_t)ldp16 : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const float pv = v[i] * inv;
        if (c < M) p[c] = pv;
        if (p16 && c < pad16) p16[c] = __float2bfloat16_rn(c < M ? pv : 0.f);
    }
    if (p16)
        for (int c = 32 * NV + lane; c < pad16; c += 32) p16[c] = __float2bfloat16_rn(0.f);
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* in, int ldi, int in_slot, const int32_t* idx,
                                                          float* out, int ldo, int B, int slot, const int32_t* len, int E) {
    int b, t;
    if (!row_of_warp(B, slot, len, b, t)) return;
    const int lane = threadIdx.x & 31;
    const size_t row = (size_t)b * slot + t;
    const size_t src = (size_t)b * in_slot + idx[row];
    for (int c = lane; c < E; c += 32) out[row * ldo + c] = in[src * ldi + c];
}

// Input staging: features stored channel-major on disk ((D, T) arrays that the reference transposes on the host,
// utils/dataset.py:12-21) are copied to the device as they are and transposed here into the packed [B][slot][D] rows
// (optionally cast to bf16).  src video b starts at src + b * src_bstride and is a dense [D][len[b]] fp32 array.
// 32 x 32 tiles through shared memory (padded: conflict-free), coalesced on both sides.
template <typename TY>
__global__ void __launch_bounds__(256) transpose_rows_kernel(const float* __restrict__ src, long long src_bstride, TY* __restrict__ dst,
                                                             int ldd, int D, int slot, const int32_t* __restrict__ len) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, T = min(len[b], slot);
    const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    if (t0 >= T) return;
    const float* s = src + (size_t)b * (size_t)src_bstride;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int d = d0 + ty + i, t = t0 + tx;
        tile[ty + i][tx] = (d < D && t < T) ? s[(size_t)d * T + t] : 0.f;
    }
    __syncthreads();
    TY* o = dst + (size_t)b * slot * ldd;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int t = t0 + ty + i, d = d0 + tx;
        if (t < T && d < D) o[(size_t)t * ldd + d] = from_f<TY>(tile[tx][ty + i]);
    }
}

static inline unsigned row_grid(int B, int slot) { return (unsigned)(((long)B * slot + ROWS_PER_CTA - 1) / ROWS_PER_CTA); }

}  // namespace factk

using namespace factk;

extern "C" int factk_softmax_splice(void* X, int dtype, int B, int slot, const int32_t* len, int ld, int H, int C,
                                    float* clogit_out, int32_t* pred_out, void* stream) {
    FACTK_REQUIRE(X && clogit_out && B > 0 && slot > 0 && C > 0 && C <= H && H <= ld, "factk_softmax_splice: bad args");
    const cudaStream_t st = (cudaStream_t)stream;
    const int nv = (C + 31) / 32;
#define SPLICE(NV_)                                                                                                              \
    (dtype == FACTK_BF16 ? splice_small_kernel<__nv_bfloat16, NV_><<<row_grid(B, slot), 256, 0, st>>>(                          \
                               reinterpret_cast<__nv_bfloat16*>(X), B, slot, len, ld, H, C, clogit_out, pred_out)               \
                         : splice_small_kernel<float, NV_><<<row_grid(B, slot), 256, 0, st>>>(reinterpret_cast<float*>(X), B,    \
                                                                                                slot, len, ld, H, C, clogit_out, \
                                                                                                pred_out))
    // frame-level launches of the shipped class counts: four rows per warp (splice_multi_kernel)
    if (dtype == FACTK_BF16 && nv == 3 && (slot % 4) == 0 && (long)B * slot >= 32768) {
        const unsigned grid4 = (unsigned)(((long)B * slot / 4 + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
        splice_multi_kernel<__nv_bfloat16, 3, 4><<<grid4, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(X), B, slot, len, ld, H, C, clogit_out, pred_out);
        return check_launch("factk_softmax_splice");
    }
    if (nv == 1) SPLICE(1);
    else if (nv == 2) SPLICE(2);
    else if (nv == 3) SPLICE(3);
    else if (nv == 4) SPLICE(4);
    else softmax_splice_kernel<<<row_grid(B, slot), 256, 0, st>>>(X, dtype, B, slot, len, ld, H, C, clogit_out, pred_out);
#undef SPLICE
    return check_launch("factk_softmax_splice");
}

extern "C" int factk_layernorm(const void* X, int x_dtype, int ldx, const void* R, int r_dtype, int ldr, const float* w,
                               const float* b, float eps, int relu, void* Y, int y_dtype, int ldy, int B, int slot,
                               const int32_t* len, int E, void* stream) {
    FACTK_REQUIRE(X && Y && w && b && B > 0 && slot > 0 && E > 0, "factk_layernorm: bad args");
    const cudaStream_t st = (cudaStream_t)stream;
    const bool r_ok = R == nullptr || r_dtype == FACTK_F32;
#define LN_GO(TX, TY, NV_)                                                                                                       \
    layernorm_small_kernel<TX, TY, NV_><<<row_grid(B, slot), 256, 0, st>>>(reinterpret_cast<const TX*>(X), ldx,                  \
                                                                           reinterpret_cast<const float*>(R), ldr, w, b, eps,    \
                                                                           relu, reinterpret_cast<TY*>(Y), ldy, B, slot, len)
#define LN_NV(TX, TY)                                                                                                            \
    do {                                                                                                                         \
        if (E == 128) { LN_GO(TX, TY, 4); return check_launch("factk_layernorm"); }                                              \
        if (E == 256) { LN_GO(TX, TY, 8); return check_launch("factk_layernorm"); }                                              \
        if (E == 512) { LN_GO(TX, TY, 16); return check_launch("factk_layernorm"); }                                             \
    } while (0)
    if (r_ok && x_dtype == FACTK_F32 && y_dtype == FACTK_F32) LN_NV(float, float);
    if (r_ok && x_dtype == FACTK_BF16 && y_dtype == FACTK_BF16) LN_NV(__nv_bfloat16, __nv_bfloat16);
    if (r_ok && x_dtype == FACTK_F32 && y_dtype == FACTK_BF16) LN_NV(float, __nv_bfloat16);
#undef LN_NV
#undef LN_GO
    layernorm_kernel<<<row_grid(B, slot), 256, 0, st>>>(X, x_dtype, ldx, R, r_dtype, ldr, w, b, eps, relu, Y, y_dtype, ldy, B, slot, len, E);
    return check_launch("factk_layernorm");
}

extern "C" int factk_l2norm(const void* X, int x_dtype, int ldx, void* Y, int y_dtype, int ldy, int B, int slot,
                            const int32_t* len, int E, float eps, void* stream) {
    FACTK_REQUIRE(X && Y && B > 0 && slot > 0 && E > 0, "factk_l2norm: bad args");
    l2norm_kernel<<<row_grid(B, slot), 256, 0, (cudaStream_t)stream>>>(X, x_dtype, ldx, Y, y_dtype, ldy, B, slot, len, E, eps);
    return check_launch("factk_l2norm");
}

extern "C" int factk_row_softmax(const float* L, int ldl, float* P, int ldp, int B, int slot, const int32_t* len, int M,
                                 float scale, void* P16, int ldp16, int pad16, void* stream) {
    FACTK_REQUIRE(L && P && B > 0 && slot > 0 && M > 0, "factk_row_softmax: bad args");
    FACTK_REQUIRE(P16 == nullptr || (pad16 >= M && pad16 <= ldp16), "factk_row_softmax: bad bf16 copy shape");
    const cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* p16 = reinterpret_cast<__nv_bfloat16*>(P16);
    const int nv = (M + 31) / 32;
    if (nv == 1) row_softmax_small_kernel<1><<<row_grid(B, slot), 256, 0, st>>>(L, ldl, P, ldp, B, slot, len, M, scale, p16, ldp16, pad16);
    else if (nv == 2) row_softmax_small_kernel<2><<<row_grid(B, slot), 256, 0, st>>>(L, ldl, P, ldp, B, slot, len, M, scale, p16, ldp16, pad16);
    else if (nv == 3) row_softmax_small_kernel<3><<<row_grid(B, slot), 256, 0, st>>>(L, ldl, P, ldp, B, slot, len, M, scale, p16, ldp16, pad16);
    else if (nv == 4) row_softmax_small_kernel<4><<<row_grid(B, slot), 256, 0, st>>>(L, ldl, P, ldp, B, slot, len, M, scale, p16, ldp16, pad16);
    else row_softmax_kernel<<<row_grid(B, slot), 256, 0, st>>>(L, ldl, P, ldp, B, slot, len, M, scale, p16, ldp16, pad16);
    return check_launch("factk_row_softmax");
}

extern "C" int factk_transpose_rows(const float* src, long long src_bstride, void* dst, int dst_dtype, int ldd, int B, int slot,
                                    int D, const int32_t* len, void* stream) {
    FACTK_REQUIRE(src && dst && len && B > 0 && slot > 0 && D > 0 && ldd >= D, "factk_transpose_rows: bad args");
    const dim3 grid((slot + 31) / 32, (D + 31) / 32, B);
    if (dst_dtype == FACTK_BF16)
        transpose_rows_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_bstride, reinterpret_cast<__nv_bfloat16*>(dst), ldd, D, slot, len);
    else
        transpose_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_bstride, reinterpret_cast<float*>(dst), ldd, D, slot, len);
    return check_launch("factk_transpose_rows");
}

extern "C" int factk_gather_rows(const float* in, int ldi, int in_slot, const int32_t* idx, float* out, int ldo, int B,
                                 int slot, const int32_t* len, int E, void* stream) {
    FACTK_REQUIRE(in && idx && out && B > 0 && slot > 0 && E > 0, "factk_gather_rows: bad args");
    gather_rows_kernel<<<row_grid(B, slot), 256, 0, (cudaStream_t)stream>>>(in, ldi, in_slot, idx, out, ldo, B, slot, len, E);
    return check_launch("factk_gather_rows");
}
