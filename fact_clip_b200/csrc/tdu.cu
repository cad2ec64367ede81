// Temporal down/up-sampling entirely on device (no host round trip):
//  * tdu_segment  : run-length segmentation of the per-frame argmax (utils/utils.py:25-48,
//                   models/basic.py:597-607, models/blocks.py:454) -- block scan per video.
//  * segment_mean : deterministic mean over each contiguous run (models/basic.py:615-625).
//  * gru_bidir    : the bidirectional GRU recurrence over segments (models/blocks.py:401,432) as a
//                   persistent kernel on 4-CTA clusters: each CTA keeps the W_hh rows of a quarter of the
//                   hidden units resident in shared memory and the new hidden state is exchanged through
//                   distributed shared memory once per step.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace factk {

__global__ void __launch_bounds__(1024) tdu_segment_kernel(const int32_t* __restrict__ pred, int slot,
                                                           const int32_t* __restrict__ len, int32_t* seg_label,
                                                           int32_t* seg_start, int32_t* seg_len, int32_t* seg_center,
                                                           int32_t* nseg) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int b = blockIdx.x;
    const int T = len ? min(len[b], slot) : slot;
    const size_t base = (size_t)b * slot;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += 1024) {
        const int t = t0 + tid;
        int flag = 0;
        if (t < T) flag = (t == 0) || (pred[base + t] != pred[base + t - 1]);
        int v = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) warp_tot[w] = v;
        __syncthreads();
        if (w == 0) {
            int x = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += n;
            }
            warp_tot[lane] = x;   // inclusive totals
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = carry + v + (w > 0 ? warp_tot[w - 1] : 0);
        if (t < T) {
            seg_label[base + t] = incl - 1;
            if (flag) seg_start[base + incl - 1] = t;
        }
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
    const int S = carry_s;
    if (tid == 0) nseg[b] = S;
    for (int s = tid; s < S; s += 1024) {
        const int st = seg_start[base + s];
        const int en = (s + 1 < S) ? seg_start[base + s + 1] : T;   // exclusive
        seg_len[base + s] = en - st;
        seg_center[base + s] = (st + en - 1) / 2;
    }
}

__global__ void __launch_bounds__(128) segment_mean_kernel(const void* __restrict__ X, int x_dtype, int ldx, void* seg,
                                                           int s_dtype, int lds, const int32_t* __restrict__ seg_start,
                                                           const int32_t* __restrict__ seg_len,
                                                           const int32_t* __restrict__ nseg, int slot, int E) {
    const int s = blockIdx.x, b = blockIdx.y;
    if (s >= nseg[b]) return;
    const size_t base = (size_t)b * slot;
    const int st = seg_start[base + s], n = seg_len[base + s];
    const float inv = 1.f / (float)n;
    const bool vec = ((reinterpret_cast<uintptr_t>(X) & 15u) == 0) && ((ldx & 3) == 0) && ((E & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(seg) & 15u) == 0) && ((lds & 3) == 0);
    if (vec) {
        for (int c = threadIdx.x * 4; c < E; c += blockDim.x * 4) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < n; ++r) {
                const float4 v = ld_vec4(X, x_dtype, (base + st + r) * (size_t)ldx + c);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            st_vec4(seg, s_dtype, (base + s) * (size_t)lds + c, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
        }
    } else {
        for (int c = threadIdx.x; c < E; c += blockDim.x) {
            float a = 0.f;
            for (int r = 0; r < n; ++r) a += ld_elem(X, x_dtype, (base + st + r) * (size_t)ldx + c);
            st_elem(seg, s_dtype, (base + s) * (size_t)lds + c, a * inv);
        }
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int GRU_CS = 4;   // CTAs per cluster
constexpr int GRU_NB = 4;   // chains (videos) advanced together by one cluster

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __cluster_dims__(GRU_CS, 1, 1)
gru_cluster_kernel(const float* __restrict__ gi, const float* __restrict__ whh_f, const float* __restrict__ bhh_f,
                   const float* __restrict__ whh_b, const float* __restrict__ bhh_b, int Hh, void* out, int o_dtype,
                   int ldo, int relu, int B, int slot, const int32_t* __restrict__ nseg) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / GRU_CS;
    const int dir = cid & 1, grp = cid >> 1;
    const int U = Hh / GRU_CS, R = 3 * U;
    extern __shared__ __align__(16) float sm[];
    float* Wt = sm;                                  // [Hh][R]   transposed slice of W_hh
    float* hb = Wt + (size_t)Hh * R;                 // [2][NB][Hh] double-buffered hidden state
    float* gx = hb + 2 * GRU_NB * Hh;                // [NB][2][U]  sigmoid(r), sigmoid(z)
    const float* whh = dir ? whh_b : whh_f;
    const float* bhh = dir ? bhh_b : bhh_f;
    const int tid = threadIdx.x;

    for (int i = tid; i < Hh * R; i += blockDim.x) {
        const int k = i / R, j = i % R, g = j / U, u = j % U;
        Wt[i] = whh[((size_t)g * Hh + rank * U + u) * Hh + k];
    }
    for (int i = tid; i < 2 * GRU_NB * Hh; i += blockDim.x) hb[i] = 0.f;

    int S[GRU_NB], vb[GRU_NB], maxS = 0;
#pragma unroll
    for (int nb = 0; nb < GRU_NB; ++nb) {
        vb[nb] = grp * GRU_NB + nb;
        S[nb] = (vb[nb] < B) ? min(nseg[vb[nb]], slot) : 0;
        maxS = max(maxS, S[nb]);
    }
    const bool active = tid < R;
    const int g = active ? tid / U : 0, u = active ? tid % U : 0, unit = rank * U + u;
    const float bias = active ? bhh[g * Hh + unit] : 0.f;
    const size_t gstride = (size_t)6 * Hh;

    float gin[GRU_NB];
    auto load_gi = [&](int t) {
#pragma unroll
        for (int nb = 0; nb < GRU_NB; ++nb) {
            gin[nb] = 0.f;
            if (active && t < S[nb]) {
                const int s = dir ? S[nb] - 1 - t : t;
                gin[nb] = gi[((size_t)vb[nb] * slot + s) * gstride + (size_t)dir * 3 * Hh + (size_t)g * Hh + unit];
            }
        }
    };
    load_gi(0);
    __syncthreads();
    cluster.sync();

    for (int t = 0; t < maxS; ++t) {
        const int cur = t & 1;
        const float* h = hb + cur * GRU_NB * Hh;
        float acc[GRU_NB];
#pragma unroll
        for (int nb = 0; nb < GRU_NB; ++nb) acc[nb] = bias;
        if (active) {
            for (int k = 0; k < Hh; k += 4) {
                const float w0 = Wt[(size_t)(k + 0) * R + tid], w1 = Wt[(size_t)(k + 1) * R + tid];
                const float w2 = Wt[(size_t)(k + 2) * R + tid], w3 = Wt[(size_t)(k + 3) * R + tid];
#pragma unroll
                for (int nb = 0; nb < GRU_NB; ++nb) {
                    const float4 h4 = *reinterpret_cast<const float4*>(&h[nb * Hh + k]);
                    acc[nb] = fmaf(w0, h4.x, acc[nb]); acc[nb] = fmaf(w1, h4.y, acc[nb]);
                    acc[nb] = fmaf(w2, h4.z, acc[nb]); acc[nb] = fmaf(w3, h4.w, acc[nb]);
                }
            }
        }
        float gcur[GRU_NB];
#pragma unroll
        for (int nb = 0; nb < GRU_NB; ++nb) gcur[nb] = gin[nb];
        if (t + 1 < maxS) load_gi(t + 1);
        if (active && g < 2) {
#pragma unroll
            for (int nb = 0; nb < GRU_NB; ++nb) gx[(nb * 2 + g) * U + u] = sigmoidf_(gcur[nb] + acc[nb]);
        }
        __syncthreads();
        if (active && g == 2) {
#pragma unroll
            for (int nb = 0; nb < GRU_NB; ++nb) {
                if (t < S[nb]) {
                    const float r = gx[(nb * 2 + 0) * U + u], z = gx[(nb * 2 + 1) * U + u];
                    const float n = tanhf(gcur[nb] + r * acc[nb]);
                    const float hn = (1.f - z) * n + z * h[nb * Hh + unit];
#pragma unroll
                    for (int rr = 0; rr < GRU_CS; ++rr) {
                        float* dst = cluster.map_shared_rank(hb, rr);
                        dst[(cur ^ 1) * GRU_NB * Hh + nb * Hh + unit] = hn;
                    }
                    const int s = dir ? S[nb] - 1 - t : t;
                    st_elem(out, o_dtype, ((size_t)vb[nb] * slot + s) * (size_t)ldo + (size_t)dir * Hh + unit,
                            relu ? fmaxf(hn, 0.f) : hn);
                }
            }
        }
        cluster.sync();
    }
}

}  // namespace factk

using namespace factk;

extern "C" int factk_tdu_segment(const int32_t* pred, int B, int slot, const int32_t* len, int32_t* seg_label,
                                 int32_t* seg_start, int32_t* seg_len, int32_t* seg_center, int32_t* nseg, void* stream) {
    FACTK_REQUIRE(pred && seg_label && seg_start && seg_len && seg_center && nseg && B > 0 && slot > 0,
                  "factk_tdu_segment: bad args");
    tdu_segment_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(pred, slot, len, seg_label, seg_start, seg_len, seg_center, nseg);
    return check_launch("factk_tdu_segment");
}

extern "C" int factk_segment_mean(const void* X, int x_dtype, int ldx, void* seg, int s_dtype, int lds,
                                  const int32_t* seg_start, const int32_t* seg_len, const int32_t* nseg, int B, int slot,
                                  int E, void* stream) {
    FACTK_REQUIRE(X && seg && seg_start && seg_len && nseg && B > 0 && slot > 0 && E > 0, "factk_segment_mean: bad args");
    FACTK_REQUIRE(B <= 65535, "factk_segment_mean: B too large");
    segment_mean_kernel<<<dim3(slot, B), 128, 0, (cudaStream_t)stream>>>(X, x_dtype, ldx, seg, s_dtype, lds, seg_start,
                                                                          seg_len, nseg, slot, E);
    return check_launch("factk_segment_mean");
}

extern "C" int factk_gru_bidir(const float* gi, const float* w_hh_f, const float* b_hh_f, const float* w_hh_b,
                               const float* b_hh_b, int Hh, void* out, int o_dtype, int ldo, int relu, int B, int slot,
                               const int32_t* nseg, void* stream) {
    FACTK_REQUIRE(gi && w_hh_f && b_hh_f && w_hh_b && b_hh_b && out && nseg && B > 0 && slot > 0, "factk_gru_bidir: bad args");
    FACTK_REQUIRE(Hh > 0 && Hh % (4 * GRU_CS) == 0, "factk_gru_bidir: hidden size %d must be a multiple of %d", Hh, 4 * GRU_CS);
    const int U = Hh / GRU_CS, R = 3 * U;
    const size_t smem = ((size_t)Hh * R + 2 * GRU_NB * Hh + GRU_NB * 2 * U) * sizeof(float);
    FACTK_REQUIRE(smem <= 227 * 1024 && R <= 1024, "factk_gru_bidir: hidden size %d does not fit one cluster (%zu B smem)", Hh, smem);
    cudaError_t e = cudaFuncSetAttribute(gru_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("factk_gru_bidir: smem attr: %s", cudaGetErrorString(e)); return FACTK_ERR_CUDA; }
    const int groups = (B + GRU_NB - 1) / GRU_NB;
    const int threads = ((R + 31) / 32) * 32;
    gru_cluster_kernel<<<groups * 2 * GRU_CS, threads, smem, (cudaStream_t)stream>>>(gi, w_hh_f, b_hh_f, w_hh_b, b_hh_b, Hh,
                                                                                      out, o_dtype, ldo, relu, B, slot, nseg);
    return check_launch("factk_gru_bidir");
}
