// Prob fusion + argmax of the last block: Block._eval (models/blocks.py:242-261) and
// FACT_CLIP.eval_with_clip (models/blocks.py:854-887).  One warp per frame; the per-video token mask
// (argmax != null class) is rebuilt in shared memory by every CTA, so there is no host sync on
// len(action_loc) (blocks.py:252,862).
#include "common.cuh"

namespace factk {

constexpr int EV_WARPS = 8, EV_FRAMES = 256, EV_MAXM = 512;   // 256 frames per CTA: the per-CTA token statistics (M rows) are amortised over 4x more frames than with 64

__global__ void __launch_bounds__(EV_WARPS * 32) fuse_eval_kernel(const float* __restrict__ aclogit,
                                                                  const float* __restrict__ attn, int lda, int attn_slot,
                                                                  const int32_t* __restrict__ seg_label,
                                                                  const float* __restrict__ flogit, int ldf, float weight,
                                                                  int64_t* __restrict__ pred, int slot,
                                                                  const int32_t* __restrict__ len, int M, int C,
                                                                  int chunks_per_video, int f_logp) {
    __shared__ int valid[EV_MAXM];
    __shared__ int any_valid;
    const int b = blockIdx.x / chunks_per_video;
    const int t0 = (blockIdx.x % chunks_per_video) * EV_FRAMES;
    const int T = len ? min(len[b], slot) : slot;
    if (t0 >= T) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) any_valid = 0;
    __syncthreads();
    // token-side: class argmax over C+1 logits (first index on ties); valid = not the null class
    for (int m = w; m < M; m += EV_WARPS) {
        const float* row = aclogit + ((size_t)b * M + m) * (C + 1);
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c <= C; c += 32) {
            const float v = row[c];
            if (v > best) { best = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) {
            valid[m] = (bi != C);
            if (bi != C) any_valid = 1;
        }
    }
    __syncthreads();
    const bool has_action = any_valid != 0;

    for (int f = w; f < EV_FRAMES; f += EV_WARPS) {
        const int t = t0 + f;
        if (t >= T) break;
        const size_t row = (size_t)b * slot + t;
        const float* fl = flogit + row * (size_t)ldf;
        // frame-branch softmax statistics
        float fm = 0.f, fs = 1.f;        // f_logp: the row already holds log-probabilities, used as exp(.) un-normalised
        if (!f_logp) {                   // (blocks_SepVerbNoun.py:308: products of verb and noun probabilities)
            fm = -INFINITY;
            for (int c = lane; c < C; c += 32) fm = fmaxf(fm, fl[c]);
            fm = warp_max(fm);
            fs = 0.f;
            for (int c = lane; c < C; c += 32) fs += __expf(fl[c] - fm);
            fs = 1.f / warp_sum(fs);
        }

        int mstar = -1;
        float qm = 0.f, qs = 0.f;
        const float* qrow = nullptr;
        if (has_action) {
            const size_t arow = (size_t)b * attn_slot + (seg_label ? seg_label[row] : t);
            const float* ar = attn + arow * (size_t)lda;
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int m = lane; m < M; m += 32) {
                if (!valid[m]) continue;
                const float v = ar[m];
                if (v > best) { best = v; bi = m; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            mstar = bi < M ? bi : 0;     // NaN attention rows cannot index out of bounds
            qrow = aclogit + ((size_t)b * M + mstar) * (C + 1);
            qm = -INFINITY;
            for (int c = lane; c < C; c += 32) qm = fmaxf(qm, qrow[c]);
            qm = warp_max(qm);
            for (int c = lane; c < C; c += 32) qs += __expf(qrow[c] - qm);
            qs = 1.f / warp_sum(qs);
        }
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {
            float p = __expf(fl[c] - fm) * fs;
            if (has_action) p = (1.f - weight) * (__expf(qrow[c] - qm) * qs) + weight * p;
            if (p > best) { best = p; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) pred[row] = (int64_t)bi;
    }
}


// The same for C <= 32 * NV classes: the frame's class row lives in registers (one load, one exp per element) and the
// per-token softmax statistics are computed once per CTA instead of once per frame -- the generic kernel above is
// instruction-issue bound (77 % issue-active in ncu at 64 x 4096 frames).  Identical arithmetic, identical results.
template <int NV>
__global__ void __launch_bounds__(EV_WARPS * 32) fuse_eval_small_kernel(const float* __restrict__ aclogit,
                                                                        const float* __restrict__ attn, int lda, int attn_slot,
                                                                        const int32_t* __restrict__ seg_label,
                                                                        const float* __restrict__ flogit, int ldf, float weight,
                                                                        int64_t* __restrict__ pred, int slot,
                                                                        const int32_t* __restrict__ len, int M, int C,
                                                                        int chunks_per_video, int f_logp) {
    __shared__ int valid[EV_MAXM];
    __shared__ float qmx[EV_MAXM], qinv[EV_MAXM];
    __shared__ int any_valid;
    const int b = blockIdx.x / chunks_per_video;
    const int t0 = (blockIdx.x % chunks_per_video) * EV_FRAMES;
    const int T = len ? min(len[b], slot) : slot;
    if (t0 >= T) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) any_valid = 0;
    __syncthreads();
    for (int m = w; m < M; m += EV_WARPS) {
        const float* row = aclogit + ((size_t)b * M + m) * (C + 1);
        float v[NV], mx = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            v[i] = c < C ? row[c] : -INFINITY;
            if (v[i] > mx) { mx = v[i]; bi = c; }
        }
        float best = mx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) sum += __expf(v[i] - best);      // best == max over the C real classes
        sum = warp_sum(sum);
        if (lane == 0) {
            const bool is_null = row[C] > best;                        // the null class is last: it wins only strictly
            qmx[m] = best;
            qinv[m] = 1.f / sum;
            valid[m] = !is_null;
            if (!is_null) any_valid = 1;
        }
    }
    __syncthreads();
    const bool has_action = any_valid != 0;

    for (int f = w; f < EV_FRAMES; f += EV_WARPS) {
        const int t = t0 + f;
        if (t >= T) break;
        const size_t row = (size_t)b * slot + t;
        const float* fl = flogit + row * (size_t)ldf;
        float fe[NV], fm = -INFINITY;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            fe[i] = c < C ? fl[c] : -INFINITY;
            fm = fmaxf(fm, fe[i]);
        }
        float fs = 1.f;
        if (!f_logp) {
            fm = warp_max(fm);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) { fe[i] = __expf(fe[i] - fm); s += fe[i]; }
            fs = 1.f / warp_sum(s);
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) fe[i] = __expf(fe[i]);
        }
        const float* qrow = nullptr;
        float qm = 0.f, qs = 0.f;
        if (has_action) {
            const size_t arow = (size_t)b * attn_slot + (seg_label ? seg_label[row] : t);
            const float* ar = attn + arow * (size_t)lda;
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int m = lane; m < M; m += 32) {
                if (!valid[m]) continue;
                const float v = ar[m];
                if (v > best) { best = v; bi = m; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            const int mstar = bi < M ? bi : 0;     // NaN attention rows cannot index out of bounds
            qrow = aclogit + ((size_t)b * M + mstar) * (C + 1);
            qm = qmx[mstar];
            qs = qinv[mstar];
        }
        float best = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                float p = fe[i] * fs;
                if (has_action) p = (1.f - weight) * (__expf(qrow[c] - qm) * qs) + weight * p;
                if (p > best) { best = p; bi = c; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) pred[row] = (int64_t)bi;
    }
}

// Block._eval_w_transcript (models/blocks.py:263-275), FACT.trans models: the tokens ARE the video's transcript, so the
// fusion runs over the N transcript positions: prob[n] = (1-w) softmax_n(attn[t, :N]) + w softmax_c(frame logits)[transcript[n]],
// pred[t] = transcript[argmax_n prob].  One warp per frame.
__global__ void __launch_bounds__(EV_WARPS * 32) fuse_eval_transcript_kernel(const float* __restrict__ attn, int lda, int attn_slot,
                                                                             const int32_t* __restrict__ seg_label,
                                                                             const float* __restrict__ flogit, int ldf, float weight,
                                                                             const int32_t* __restrict__ transcript, int ldt,
                                                                             const int32_t* __restrict__ ntr, int64_t* __restrict__ pred,
                                                                             int slot, const int32_t* __restrict__ len, int C,
                                                                             int chunks_per_video) {
    const int b = blockIdx.x / chunks_per_video;
    const int t0 = (blockIdx.x % chunks_per_video) * EV_FRAMES;
    const int T = len ? min(len[b], slot) : slot;
    if (t0 >= T) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int N = ntr[b];
    const int32_t* tr = transcript + (size_t)b * ldt;
    for (int f = w; f < EV_FRAMES; f += EV_WARPS) {
        const int t = t0 + f;
        if (t >= T) break;
        const size_t row = (size_t)b * slot + t;
        const float* fl = flogit + row * (size_t)ldf;
        float fm = -INFINITY;
        for (int c = lane; c < C; c += 32) fm = fmaxf(fm, fl[c]);
        fm = warp_max(fm);
        float fs = 0.f;
        for (int c = lane; c < C; c += 32) fs += __expf(fl[c] - fm);
        fs = 1.f / warp_sum(fs);
        const float* ar = attn + ((size_t)b * attn_slot + (seg_label ? seg_label[row] : t)) * (size_t)lda;
        float am = -INFINITY;
        for (int n = lane; n < N; n += 32) am = fmaxf(am, ar[n]);
        am = warp_max(am);
        float as = 0.f;
        for (int n = lane; n < N; n += 32) as += __expf(ar[n] - am);
        as = 1.f / warp_sum(as);
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int n = lane; n < N; n += 32) {
            const float p = (1.f - weight) * (__expf(ar[n] - am) * as) + weight * (__expf(fl[tr[n]] - fm) * fs);
            if (p > best) { best = p; bi = n; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) pred[row] = (int64_t)tr[bi < N ? bi : 0];
    }
}

// Token initialisation of FACT.trans models (models/blocks.py:74-79): out[n, :] = action_embed[transcript[n], :] + pe[n, :].
__global__ void embed_tokens_kernel(const float* __restrict__ embed, int lde, const int32_t* __restrict__ transcript,
                                    const float* __restrict__ pe, int ldpe, float* __restrict__ out, int ldo, int N, int A) {
    const int n = blockIdx.x;
    if (n >= N) return;
    const int cls = transcript[n];
    for (int a = threadIdx.x; a < A; a += blockDim.x) out[(size_t)n * ldo + a] = embed[(size_t)cls * lde + a] + pe[(size_t)n * ldpe + a];
}

}  // namespace factk

using namespace factk;

extern "C" int factk_fuse_eval(const float* action_clogit, const float* attn, int lda, int attn_slot,
                               const int32_t* seg_label, const float* flogit, int ldf, float weight, int64_t* pred, int B,
                               int slot, const int32_t* len, int M, int C, int f_logp, void* stream) {
    FACTK_REQUIRE((M == 0 || action_clogit) && flogit && pred && B > 0 && slot > 0 && C > 0, "factk_fuse_eval: bad args");
    FACTK_REQUIRE(M >= 0 && M <= EV_MAXM, "factk_fuse_eval: at most %d tokens", EV_MAXM);
    FACTK_REQUIRE(M == 0 || attn != nullptr, "factk_fuse_eval: attention required when M > 0");
    const int cpv = (slot + EV_FRAMES - 1) / EV_FRAMES;
    const cudaStream_t st = (cudaStream_t)stream;
    const int nv = (C + 31) / 32;
#define FEV(NV_) fuse_eval_small_kernel<NV_><<<(unsigned)(cpv * B), EV_WARPS * 32, 0, st>>>(                                   \
        action_clogit, attn, lda, attn_slot, seg_label, flogit, ldf, weight, pred, slot, len, M, C, cpv, f_logp)
    if (M > 0 && nv == 1) FEV(1);
    else if (M > 0 && nv == 2) FEV(2);
    else if (M > 0 && nv == 3) FEV(3);
    else if (M > 0 && nv == 4) FEV(4);
    else
        fuse_eval_kernel<<<(unsigned)(cpv * B), EV_WARPS * 32, 0, st>>>(action_clogit, attn, lda, attn_slot, seg_label, flogit, ldf,
                                                                        weight, pred, slot, len, M, C, cpv, f_logp);
#undef FEV
    return check_launch("factk_fuse_eval");
}

extern "C" int factk_fuse_eval_transcript(const float* attn, int lda, int attn_slot, const int32_t* seg_label, const float* flogit,
                                          int ldf, float weight, const int32_t* transcript, int ldt, const int32_t* ntr,
                                          int64_t* pred, int B, int slot, const int32_t* len, int C, void* stream) {
    FACTK_REQUIRE(attn && flogit && transcript && ntr && pred && B > 0 && slot > 0 && C > 0, "factk_fuse_eval_transcript: bad args");
    const int cpv = (slot + EV_FRAMES - 1) / EV_FRAMES;
    fuse_eval_transcript_kernel<<<(unsigned)(cpv * B), EV_WARPS * 32, 0, (cudaStream_t)stream>>>(
        attn, lda, attn_slot, seg_label, flogit, ldf, weight, transcript, ldt, ntr, pred, slot, len, C, cpv);
    return check_launch("factk_fuse_eval_transcript");
}

extern "C" int factk_embed_tokens(const float* embed, int lde, const int32_t* transcript, const float* pe, int ldpe, float* out,
                                  int ldo, int N, int A, void* stream) {
    FACTK_REQUIRE(embed && transcript && pe && out && N > 0 && A > 0, "factk_embed_tokens: bad args");
    embed_tokens_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(embed, lde, transcript, pe, ldpe, out, ldo, N, A);
    return check_launch("factk_embed_tokens");
}
