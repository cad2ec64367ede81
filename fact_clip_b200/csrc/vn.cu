// Epic-Kitchens verb/noun heads (reference models/blocks_SepVerbNoun.py): every feature carries two class heads
// (verbs | nouns) that are softmaxed separately, and an action class a is the pair (vids[a], nids[a]).
//  * vn_splice_kernel  : Block.process_feature (:229-234) -- two softmaxes spliced back into the feature row, raw logits
//                        stashed -- fused with the segmentation argmax of temporal_downsample (:281-290): the action
//                        whose verb-prob x noun-prob product is largest (first index on ties)
//  * vn_combine_kernel : combine_verb_noun_to_action(apply_log=True) (:188-226): action log-probabilities
// One warp per row; the row's probabilities are staged in shared memory because the action table indexes them freely.
#include "common.cuh"

namespace factk {

constexpr int VN_WARPS = 8;

__device__ __forceinline__ void head_stats(const float* x, int n, int lane, float& mx, float& inv) {
    mx = -INFINITY;
    for (int c = lane; c < n; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < n; c += 32) s += __expf(x[c] - mx);
    inv = 1.f / warp_sum(s);
}

__global__ void __launch_bounds__(VN_WARPS * 32) vn_splice_kernel(void* X, int dtype, int B, int slot,
                                                                  const int32_t* __restrict__ len, int ld, int H, int n1, int n2,
                                                                  float* __restrict__ clogit, const int32_t* __restrict__ vids,
                                                                  const int32_t* __restrict__ nids, int nact,
                                                                  int32_t* __restrict__ pred) {
    extern __shared__ float smv[];       // [VN_WARPS][n1 + n2] logits, then probabilities
    const long r = (long)blockIdx.x * VN_WARPS + (threadIdx.x >> 5);
    if (r >= (long)B * slot) return;
    const int b = (int)(r / slot), t = (int)(r % slot), lane = threadIdx.x & 31, C = n1 + n2;
    if (t >= (len ? len[b] : slot)) return;
    float* p = smv + (threadIdx.x >> 5) * C;
    const size_t base = (size_t)r * ld + (H - C);
    for (int c = lane; c < C; c += 32) {
        const float l = ld_elem(X, dtype, base + c);
        p[c] = l;
        clogit[(size_t)r * C + c] = l;
    }
    __syncwarp();
    float m1, i1, m2, i2;
    head_stats(p, n1, lane, m1, i1);
    head_stats(p + n1, n2, lane, m2, i2);
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
        const float q = c < n1 ? __expf(p[c] - m1) * i1 : __expf(p[c] - m2) * i2;
        p[c] = q;
        st_elem(X, dtype, base + c, q);
    }
    __syncwarp();
    if (!pred) return;
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int a = lane; a < nact; a += 32) {
        const float q = p[vids[a]] * p[n1 + nids[a]];
        if (q > best) { best = q; bi = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) pred[r] = bi;
}

__global__ void __launch_bounds__(VN_WARPS * 32) vn_combine_kernel(const float* __restrict__ clogit, int ldc, int k1, int k2,
                                                                   const int32_t* __restrict__ vids, const int32_t* __restrict__ nids,
                                                                   int nact, int with_null, float* __restrict__ out, int ldo, int B,
                                                                   int slot, const int32_t* __restrict__ len) {
    extern __shared__ float smv[];       // [VN_WARPS][k1 + k2] log-probabilities
    const long r = (long)blockIdx.x * VN_WARPS + (threadIdx.x >> 5);
    if (r >= (long)B * slot) return;
    const int b = (int)(r / slot), t = (int)(r % slot), lane = threadIdx.x & 31;
    if (t >= (len ? len[b] : slot)) return;
    float* p = smv + (threadIdx.x >> 5) * (k1 + k2);
    const float* x = clogit + (size_t)r * ldc;
    float mx = -INFINITY;
    for (int c = lane; c < k1; c += 32) mx = fmaxf(mx, x[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < k1; c += 32) s += expf(x[c] - mx);
    const float l1 = mx + logf(warp_sum(s));
    mx = -INFINITY;
    for (int c = lane; c < k2; c += 32) mx = fmaxf(mx, x[k1 + c]);
    mx = warp_max(mx);
    s = 0.f;
    for (int c = lane; c < k2; c += 32) s += expf(x[k1 + c] - mx);
    const float l2 = mx + logf(warp_sum(s));
    for (int c = lane; c < k1 + k2; c += 32) p[c] = x[c] - (c < k1 ? l1 : l2);
    __syncwarp();
    float* o = out + (size_t)r * ldo;
    for (int a = lane; a < nact; a += 32) o[a] = p[vids[a]] + p[k1 + nids[a]];
    if (with_null && lane == 0) o[nact] = p[k1 - 1] + p[k1 + k2 - 1];
}

}  // namespace factk

using namespace factk;

extern "C" int factk_vn_splice(void* X, int dtype, int B, int slot, const int32_t* len, int ld, int H, int n1, int n2,
                               float* clogit_out, const int32_t* vids, const int32_t* nids, int nact, int32_t* pred_out,
                               void* stream) {
    FACTK_REQUIRE(X && clogit_out && B > 0 && slot > 0 && n1 > 0 && n2 > 0 && n1 + n2 <= H && H <= ld, "factk_vn_splice: bad args");
    FACTK_REQUIRE(pred_out == nullptr || (vids && nids && nact > 0), "factk_vn_splice: the segmentation argmax needs the action table");
    const size_t smem = (size_t)VN_WARPS * (n1 + n2) * sizeof(float);
    FACTK_REQUIRE(smem <= 48 * 1024, "factk_vn_splice: %d classes do not fit the staging buffer", n1 + n2);
    const long rows = (long)B * slot;
    vn_splice_kernel<<<(unsigned)((rows + VN_WARPS - 1) / VN_WARPS), VN_WARPS * 32, smem, (cudaStream_t)stream>>>(
        X, dtype, B, slot, len, ld, H, n1, n2, clogit_out, vids, nids, nact, pred_out);
    return check_launch("factk_vn_splice");
}

extern "C" int factk_vn_combine(const float* clogit, int ldc, int k1, int k2, const int32_t* vids, const int32_t* nids, int nact,
                                int with_null, float* out, int ldo, int B, int slot, const int32_t* len, void* stream) {
    FACTK_REQUIRE(clogit && vids && nids && out && B > 0 && slot > 0 && k1 > 0 && k2 > 0 && nact > 0 && ldo >= nact + (with_null ? 1 : 0),
                  "factk_vn_combine: bad args");
    const size_t smem = (size_t)VN_WARPS * (k1 + k2) * sizeof(float);
    FACTK_REQUIRE(smem <= 48 * 1024, "factk_vn_combine: %d classes do not fit the staging buffer", k1 + k2);
    const long rows = (long)B * slot;
    vn_combine_kernel<<<(unsigned)((rows + VN_WARPS - 1) / VN_WARPS), VN_WARPS * 32, smem, (cudaStream_t)stream>>>(
        clogit, ldc, k1, k2, vids, nids, nact, with_null, out, ldo, B, slot, len);
    return check_launch("factk_vn_combine");
}
