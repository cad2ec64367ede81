// Training-loss VALUE of FACT / FACT_CLIP on device (reference models/loss.py + the compute_loss methods of blocks.py),
// batched over the videos of a step.  Everything here is a bandwidth-trivial reduction over tensors the forward already
// left in HBM ([T,C] class logits, [T,M] attention logits); the point is that nothing but the [M,S] matching cost and
// the final numbers ever crosses PCIe (the reference moves the (1,T,M) attention to the host and builds a (T,M,S) numpy
// temporary per video, loss.py:91-106).
//
//  * label_prep_kernel   : transcript / segment weights from the run-length coded labels, per-class frame counts
//  * gt_overlap_kernel   : overlap[s][a] = sum of a2f attention of token a over ground-truth segment s
//  * match_cost_kernel   : cost[a][s] = -pc * softmax(token logits)[a, transcript[s]] - a2fc * soft-IoU(a, s), with the
//                          union in closed form (attention <= 1  =>  min(attn + onehot, 1) = onehot ? 1 : attn)
//  * loss_pick_kernel    : sum over frames of -log_softmax(...)[target] * weight, log-softmax either over (selected)
//                          columns of the frame's row or over the rows of the target column (col_lse given); optional
//                          frame -> segment row indirection with 1/len scaling (the "zoomed label" forms, loss.py:227-277)
//  * smooth_kernel       : clamped squared step of the log-probabilities between consecutive rows (loss.py:8-19)
//  * col_lse_kernel      : log-sum-exp over rows per column
//  * token_loss_kernel   : class-weighted cross entropy of the action tokens against their matched segments
//  * loss_combine_kernel : the per-block formulas (blocks.py:313-320, 369-382, 487-497), block mean, InfoNCE mix
//
// All sums are formed in a fixed order (per-warp, per-CTA partials, sequential combine): results are bit-reproducible.
#include "common.cuh"

namespace factk {

constexpr int LS_WARPS = 8;
constexpr int LS_CHUNK = 64;      // frames per CTA = one partial-sum slot

__device__ __forceinline__ float warp_lse(const float* row, const int32_t* cols, int K, int lane) {
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, row[cols ? cols[k] : k]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += expf(row[cols ? cols[k] : k] - mx);
    return mx + logf(warp_sum(s));
}

// Sum the LS_WARPS per-warp values in warp order and store them as this CTA's partial.
__device__ __forceinline__ void store_partial(float v, float* sm, float* dst) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < LS_WARPS; ++i) s += sm[i];
        *dst = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) loss_pick_kernel(
    const float* __restrict__ X, int ldx, int xslot, int ncol, const int32_t* __restrict__ cols, int cols_bstride,
    const int32_t* __restrict__ ncols, const int32_t* __restrict__ ridx, const int32_t* __restrict__ rlen,
    const float* __restrict__ col_lse, int ld_lse, const int32_t* __restrict__ tgt0, const int32_t* __restrict__ tmap,
    int tmap_bstride, const float* __restrict__ w, int w_bstride, float* __restrict__ part_sum,
    float* __restrict__ part_cnt, int B, int slot, const int32_t* __restrict__ len, int nchunk) {
    __shared__ float sm[LS_WARPS];
    const int b = blockIdx.y, chunk = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int T = len ? len[b] : slot;
    const int K = cols ? (ncols ? ncols[b] : ncol) : ncol;
    const int32_t* cb = cols ? cols + (size_t)b * cols_bstride : nullptr;
    float acc = 0.f, cnt = 0.f;
    for (int i = wid; i < LS_CHUNK; i += LS_WARPS) {
        const int t = chunk * LS_CHUNK + i;
        if (t >= T) break;
        int k = tgt0[(size_t)b * slot + t];
        if (tmap && k >= 0) k = tmap[(size_t)b * tmap_bstride + k];
        if (k < 0 || k >= K) continue;
        const int r = ridx ? ridx[(size_t)b * slot + t] : t;
        const float* row = X + ((size_t)b * xslot + r) * ldx;
        const int c = cb ? cb[k] : k;
        const float lse = col_lse ? col_lse[(size_t)b * ld_lse + c] : warp_lse(row, cb, K, lane);
        float v = -(row[c] - lse) * (w ? w[(size_t)b * w_bstride + k] : 1.f);
        if (rlen) v /= (float)rlen[(size_t)b * xslot + r];
        acc += v;
        cnt += 1.f;
    }
    store_partial(acc, sm, part_sum + (size_t)b * nchunk + chunk);
    if (part_cnt) store_partial(cnt, sm, part_cnt + (size_t)b * nchunk + chunk);
}

__global__ void __launch_bounds__(256) smooth_kernel(const float* __restrict__ X, int ldx, int ncol,
                                                     float* __restrict__ part, int B, int slot,
                                                     const int32_t* __restrict__ len, int nchunk, int is_logp) {
    __shared__ float sm[LS_WARPS];
    const int b = blockIdx.y, chunk = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int T = len ? len[b] : slot;
    float acc = 0.f;
    for (int i = wid; i < LS_CHUNK; i += LS_WARPS) {
        const int t = chunk * LS_CHUNK + i;
        if (t + 1 >= T) break;
        const float* r0 = X + ((size_t)b * slot + t) * ldx;
        const float* r1 = r0 + ldx;
        // is_logp: the rows already hold log-probabilities (smooth_loss(..., is_logit=False), loss.py:12-15)
        const float l0 = is_logp ? 0.f : warp_lse(r0, nullptr, ncol, lane), l1 = is_logp ? 0.f : warp_lse(r1, nullptr, ncol, lane);
        float s = 0.f;
        for (int c = lane; c < ncol; c += 32) {
            const float d = (r1[c] - l1) - (r0[c] - l0);
            s += fminf(d * d, 16.f);
        }
        acc += warp_sum(s);
    }
    store_partial(acc, sm, part + (size_t)b * nchunk + chunk);
}

// lse over rows r < nrows[b] (rows with a negative mapped target are skipped when rmask0 is given) for every column.
// With a workspace the rows are split into LC_ROWS-row chunks across CTAs (a long video at batch 1 would otherwise be ten CTAs
// walking 16384 rows) and the (max, sum) partials are combined in chunk order by col_lse_combine_kernel.
constexpr int LC_ROWS = 512;
__global__ void __launch_bounds__(256) col_lse_kernel(const float* __restrict__ X, int ldx, int xslot, int ncol,
                                                      const int32_t* __restrict__ nrows, const int32_t* __restrict__ rmask0,
                                                      const int32_t* __restrict__ rmap, int rmap_bstride,
                                                      float* __restrict__ out, int ldo, float* __restrict__ ws, int nchunk) {
    __shared__ float smx[8][33], sms[8][33];
    const int b = blockIdx.y, cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const int R = nrows[b];
    const int r0 = ws ? blockIdx.z * LC_ROWS : 0, r1 = ws ? min(R, r0 + LC_ROWS) : R;
    if (ws && r0 >= R) return;
    float mx = -INFINITY, s = 0.f;
    if (c < ncol) {
        for (int r = r0 + ry; r < r1; r += 8) {
            if (rmask0) {
                int k = rmask0[(size_t)b * xslot + r];
                if (rmap && k >= 0) k = rmap[(size_t)b * rmap_bstride + k];
                if (k < 0) continue;
            }
            const float x = X[((size_t)b * xslot + r) * ldx + c];
            if (x > mx) { s = s * expf(mx - x) + 1.f; mx = x; }
            else s += expf(x - mx);
        }
    }
    smx[ry][cx] = mx;
    sms[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < ncol) {
        float M = -INFINITY;
        for (int i = 0; i < 8; ++i) M = fmaxf(M, smx[i][cx]);
        float S = 0.f;
        for (int i = 0; i < 8; ++i) S += (smx[i][cx] == -INFINITY) ? 0.f : sms[i][cx] * expf(smx[i][cx] - M);
        if (ws) {
            float* o = ws + ((size_t)(b * nchunk + blockIdx.z) * ncol + c) * 2;
            o[0] = M; o[1] = S;
        } else {
            out[(size_t)b * ldo + c] = M + logf(S);
        }
    }
}

__global__ void col_lse_combine_kernel(const float* __restrict__ ws, int ncol, const int32_t* __restrict__ nrows, int nchunk,
                                       float* __restrict__ out, int ldo) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= ncol) return;
    const int R = nrows[b];
    float M = -INFINITY;
    for (int k = 0; k < nchunk && k * LC_ROWS < R; ++k) M = fmaxf(M, ws[((size_t)(b * nchunk + k) * ncol + c) * 2]);
    float S = 0.f;
    for (int k = 0; k < nchunk && k * LC_ROWS < R; ++k) {
        const float* p = ws + ((size_t)(b * nchunk + k) * ncol + c) * 2;
        S += (p[0] == -INFINITY) ? 0.f : p[1] * expf(p[0] - M);
    }
    out[(size_t)b * ldo + c] = M + logf(S);
}

__global__ void __launch_bounds__(256) label_prep_kernel(const int32_t* __restrict__ label, const int32_t* __restrict__ seg_start,
                                                         const int32_t* __restrict__ nseg, const float* __restrict__ cweight,
                                                         const int32_t* __restrict__ cmap, int C, int32_t* __restrict__ transcript,
                                                         float* __restrict__ sweight, int smax, float* __restrict__ inv_count,
                                                         int32_t* __restrict__ nvalid, int B, int slot,
                                                         const int32_t* __restrict__ len) {
    extern __shared__ int hist[];      // C counters
    const int b = blockIdx.x;
    const int S = nseg[b], T = len[b];
    for (int s = threadIdx.x; s < S && s < smax; s += blockDim.x) {
        const int c = label[(size_t)b * slot + seg_start[(size_t)b * slot + s]];
        transcript[(size_t)b * smax + s] = c;
        sweight[(size_t)b * smax + s] = cweight[c];
    }
    if (!inv_count) return;
    for (int c = threadIdx.x; c < C; c += blockDim.x) hist[c] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        int k = label[(size_t)b * slot + t];
        if (cmap) k = cmap[k];
        if (k >= 0) atomicAdd(&hist[k], 1);      // integer counters: order-independent
    }
    __syncthreads();
    int tot = 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) inv_count[(size_t)b * C + c] = 1.f / (float)max(hist[c], 1);
    if (threadIdx.x == 0) {
        for (int c = 0; c < C; ++c) tot += hist[c];
        nvalid[b] = tot;
    }
}

__global__ void __launch_bounds__(128) gt_overlap_kernel(const float* __restrict__ attn, int lda, int aslot,
                                                         const int32_t* __restrict__ ridx, const int32_t* __restrict__ seg_start,
                                                         const int32_t* __restrict__ seg_len, const int32_t* __restrict__ nseg,
                                                         float* __restrict__ overlap, int smax, int ldo, int M, int slot) {
    const int b = blockIdx.y, s = blockIdx.x;
    if (s >= nseg[b]) return;
    const int t0 = seg_start[(size_t)b * slot + s], n = seg_len[(size_t)b * slot + s];
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
        float acc = 0.f;
        for (int t = t0; t < t0 + n; ++t) {
            const int r = ridx ? ridx[(size_t)b * slot + t] : t;
            acc += attn[((size_t)b * aslot + r) * lda + a];
        }
        overlap[((size_t)b * smax + s) * ldo + a] = acc;
    }
}

__global__ void __launch_bounds__(256) match_cost_kernel(const float* __restrict__ aclogit, int M, int C1,
                                                         const int32_t* __restrict__ transcript, const int32_t* __restrict__ seg_len,
                                                         const int32_t* __restrict__ nseg, const float* __restrict__ overlap,
                                                         int smax, int ldo, int slot, float pc, float a2fc,
                                                         float* __restrict__ cost, int logp) {
    extern __shared__ float smf[];     // colsum[M], lse[M]
    float* colsum = smf;
    float* lse = smf + M;
    const int b = blockIdx.x, S = nseg[b], lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += overlap[((size_t)b * smax + s) * ldo + a];
        colsum[a] = acc;
    }
    for (int a = wid; a < M; a += 8) {       // logp: the rows are log-probabilities, used as exp(.) (blocks_SepVerbNoun.py:101)
        const float l = logp ? 0.f : warp_lse(aclogit + ((size_t)b * M + a) * C1, nullptr, C1, lane);
        if (lane == 0) lse[a] = l;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < M * S; i += blockDim.x) {
        const int a = i / S, s = i % S;
        float c = 0.f;
        if (pc > 0.f) c -= pc * expf(aclogit[((size_t)b * M + a) * C1 + transcript[(size_t)b * smax + s]] - lse[a]);
        if (a2fc > 0.f) {
            const float ov = overlap[((size_t)b * smax + s) * ldo + a];
            const float un = colsum[a] - ov + (float)seg_len[(size_t)b * slot + s];
            const float iou = ov / un;
            c -= a2fc * (iou == iou ? iou : 0.f);
        }
        cost[((size_t)b * M + a) * smax + s] = c;
    }
}

__global__ void __launch_bounds__(256) token_loss_kernel(const float* __restrict__ aclogit, int M, int C1,
                                                         const int32_t* __restrict__ aind, const int32_t* __restrict__ sind,
                                                         const int32_t* __restrict__ nmatch, int kmax,
                                                         const int32_t* __restrict__ transcript, int smax,
                                                         const float* __restrict__ cweight, float* __restrict__ out, int out_stride,
                                                         int logp_mean) {
    extern __shared__ float smf[];     // num[M], den[M], clabel[M]
    float* num = smf;
    float* den = smf + M;
    int* clabel = reinterpret_cast<int*>(smf + 2 * M);
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int a = threadIdx.x; a < M; a += blockDim.x) clabel[a] = C1 - 1;
    __syncthreads();
    if (threadIdx.x == 0)        // sequential: with one-to-many matches the last pair of a token wins (loss.py:203)
        for (int k = 0; k < nmatch[b]; ++k)
            clabel[aind[(size_t)b * kmax + k]] = transcript[(size_t)b * smax + sind[(size_t)b * kmax + k]];
    __syncthreads();
    for (int a = wid; a < M; a += 8) {
        const float* row = aclogit + ((size_t)b * M + a) * C1;
        const float l = logp_mean ? 0.f : warp_lse(row, nullptr, C1, lane);
        if (lane == 0) {
            const float wgt = cweight[clabel[a]];
            num[a] = -(row[clabel[a]] - l) * wgt;
            den[a] = wgt;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float n = 0.f, d = 0.f;
        for (int a = 0; a < M; ++a) { n += num[a]; d += den[a]; }
        // logp_mean: the verb/noun model's own token loss (blocks_SepVerbNoun.py:254-266) takes log-probabilities and
        // averages over the tokens instead of normalising by the class weights
        out[(size_t)b * out_stride] = logp_mean ? n / (float)M : n / d;
    }
}

struct LossPlan {
    int nb;
    int type[FACTK_LOSS_MAX_BLOCKS];      // 0 input, 1 update, 2 update with temporal down/up-sampling; 3 / 4: verb/noun input / update
};

__device__ __forceinline__ float chunk_sum(const float* ws, int term, int b, int B, int nchunk) {
    const float* p = ws + ((size_t)term * B + b) * nchunk;
    float s = 0.f;
    for (int i = 0; i < nchunk; ++i) s += p[i];
    return s;
}

// ws: [8 * nb + 3][B][nchunk] partial sums, term slots per block: 0 frame CE, 1 frame smooth, 2 token loss (value),
// 3 f2a, 4 a2f, 5 f2a smooth, 6 a2f smooth, 7 segment CE; then v2t sum, t2v sum, (unused).
__global__ void loss_combine_kernel(const float* __restrict__ ws, LossPlan plan, int B, int nchunk,
                                    const int32_t* __restrict__ len, const int32_t* __restrict__ npred, int C, int M,
                                    float sw, int use_clip, float fact_w, float con_w, int nseen,
                                    const int32_t* __restrict__ nvalid, float* __restrict__ out, int ldo) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float T = (float)len[b];
    float fact = 0.f;
    for (int i = 0; i < plan.nb; ++i) {
        const int t0 = 8 * i;
        const float fl = chunk_sum(ws, t0 + 0, b, B, nchunk) / T;
        const float sm = chunk_sum(ws, t0 + 1, b, B, nchunk) / ((T - 1.f) * (float)C);     // 0/0 = NaN for one frame, as torch.mean
        const float atk = ws[((size_t)(t0 + 2) * B + b) * nchunk];
        float l;
        if (plan.type[i] == 0) {
            l = fl + atk + sw * sm;
        } else if (plan.type[i] == 1) {
            const float f2a = chunk_sum(ws, t0 + 3, b, B, nchunk) / T, a2f = chunk_sum(ws, t0 + 4, b, B, nchunk) / T;
            const float fs = chunk_sum(ws, t0 + 5, b, B, nchunk) / ((T - 1.f) * (float)M);
            const float as = chunk_sum(ws, t0 + 6, b, B, nchunk) / ((T - 1.f) * (float)M);
            l = atk + f2a + a2f + fl + sw * (as + fs + sm);
        } else if (plan.type[i] == 2) {
            const float S = (float)npred[(size_t)i * B + b];
            const float f2a = chunk_sum(ws, t0 + 3, b, B, nchunk) / S, a2f = chunk_sum(ws, t0 + 4, b, B, nchunk) / S;
            const float seg = chunk_sum(ws, t0 + 7, b, B, nchunk) / S;
            l = (fl + seg) / 2.f + atk + f2a + a2f + sw * sm;
        } else {
            // verb/noun blocks (blocks_SepVerbNoun.py:400-413, 485-497): frame, segment and token terms enter halved
            const float S = (float)npred[(size_t)i * B + b];
            const float seg = chunk_sum(ws, t0 + 7, b, B, nchunk) / S;
            l = (fl / 2.f + seg / 2.f) / 2.f + atk / 2.f + sw * sm;
            if (plan.type[i] == 4)
                l += chunk_sum(ws, t0 + 3, b, B, nchunk) / S + chunk_sum(ws, t0 + 4, b, B, nchunk) / S;
        }
        out[(size_t)b * ldo + 4 + i] = l;
        fact += l;
    }
    fact /= (float)plan.nb;
    float total = fact, con = 0.f;
    if (use_clip && nvalid[b] > 0) {
        const float v2t = chunk_sum(ws, 8 * plan.nb + 0, b, B, nchunk) / (float)nvalid[b];
        const float t2v = chunk_sum(ws, 8 * plan.nb + 1, b, B, nchunk) / (float)nseen;
        con = (v2t + t2v) / 2.f;
        total = fact_w * fact + con_w * con;
    }
    out[(size_t)b * ldo + 0] = total;
    out[(size_t)b * ldo + 1] = fact;
    out[(size_t)b * ldo + 2] = con;
    out[(size_t)b * ldo + 3] = (use_clip && nvalid[b] > 0) ? 1.f : 0.f;
}

}  // namespace factk

using namespace factk;

extern "C" int factk_loss_pick(const float* X, int ldx, int xslot, int ncol, const int32_t* cols, int cols_bstride,
                               const int32_t* ncols, const int32_t* ridx, const int32_t* rlen, const float* col_lse,
                               int ld_lse, const int32_t* tgt0, const int32_t* tmap, int tmap_bstride, const float* w,
                               int w_bstride, float* part_sum, float* part_cnt, int B, int slot, const int32_t* len,
                               int nchunk, void* stream) {
    FACTK_REQUIRE(X && tgt0 && part_sum && B > 0 && slot > 0 && ncol > 0, "factk_loss_pick: bad args");
    FACTK_REQUIRE(nchunk * LS_CHUNK >= slot, "factk_loss_pick: %d partial slots do not cover %d frames", nchunk, slot);
    loss_pick_kernel<<<dim3(nchunk, B), 256, 0, (cudaStream_t)stream>>>(X, ldx, xslot, ncol, cols, cols_bstride, ncols, ridx, rlen,
                                                                        col_lse, ld_lse, tgt0, tmap, tmap_bstride, w, w_bstride,
                                                                        part_sum, part_cnt, B, slot, len, nchunk);
    return check_launch("factk_loss_pick");
}

extern "C" int factk_loss_smooth(const float* X, int ldx, int ncol, float* part, int B, int slot, const int32_t* len,
                                 int nchunk, int is_logp, void* stream) {
    FACTK_REQUIRE(X && part && B > 0 && slot > 0 && ncol > 0, "factk_loss_smooth: bad args");
    FACTK_REQUIRE(nchunk * LS_CHUNK >= slot, "factk_loss_smooth: %d partial slots do not cover %d frames", nchunk, slot);
    smooth_kernel<<<dim3(nchunk, B), 256, 0, (cudaStream_t)stream>>>(X, ldx, ncol, part, B, slot, len, nchunk, is_logp);
    return check_launch("factk_loss_smooth");
}

extern "C" size_t factk_col_lse_ws_floats(int B, int xslot, int ncol) {
    return xslot > 2 * factk::LC_ROWS ? (size_t)B * ((xslot + factk::LC_ROWS - 1) / factk::LC_ROWS) * ncol * 2 : 0;
}

extern "C" int factk_col_lse(const float* X, int ldx, int xslot, int ncol, const int32_t* nrows, const int32_t* rmask0,
                             const int32_t* rmap, int rmap_bstride, float* out, int ldo, int B, float* ws, void* stream) {
    FACTK_REQUIRE(X && nrows && out && B > 0 && ncol > 0 && ldo >= ncol, "factk_col_lse: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (ws != nullptr && xslot > 2 * LC_ROWS) {          // long rows tensors: chunks of rows across CTAs, fixed-order combine
        const int nchunk = (xslot + LC_ROWS - 1) / LC_ROWS;
        col_lse_kernel<<<dim3((ncol + 31) / 32, B, nchunk), 256, 0, st>>>(X, ldx, xslot, ncol, nrows, rmask0, rmap, rmap_bstride, out, ldo, ws,
                                                                         nchunk);
        col_lse_combine_kernel<<<dim3((ncol + 63) / 64, B), 64, 0, st>>>(ws, ncol, nrows, nchunk, out, ldo);
    } else {
        col_lse_kernel<<<dim3((ncol + 31) / 32, B), 256, 0, st>>>(X, ldx, xslot, ncol, nrows, rmask0, rmap, rmap_bstride, out, ldo, nullptr, 1);
    }
    return check_launch("factk_col_lse");
}

extern "C" int factk_label_prep(const int32_t* label, const int32_t* seg_start, const int32_t* nseg, const float* cweight,
                                const int32_t* cmap, int C, int32_t* transcript, float* sweight, int smax,
                                float* inv_count, int32_t* nvalid, int B, int slot, const int32_t* len, void* stream) {
    FACTK_REQUIRE(label && seg_start && nseg && cweight && transcript && sweight && len && B > 0 && smax > 0 && C > 0,
                  "factk_label_prep: bad args");
    FACTK_REQUIRE(C <= 8192 && (inv_count == nullptr) == (nvalid == nullptr), "factk_label_prep: bad class count / outputs");
    label_prep_kernel<<<B, 256, C * sizeof(int), (cudaStream_t)stream>>>(label, seg_start, nseg, cweight, cmap, C, transcript,
                                                                         sweight, smax, inv_count, nvalid, B, slot, len);
    return check_launch("factk_label_prep");
}

extern "C" int factk_match_cost(const float* attn, int lda, int aslot, const int32_t* ridx, const float* aclogit, int M,
                                int C1, const int32_t* transcript, const int32_t* seg_start, const int32_t* seg_len,
                                const int32_t* nseg, int smax, float pc, float a2fc, float* overlap, int ldo, float* cost,
                                int B, int slot, int logp, void* stream) {
    FACTK_REQUIRE(attn && aclogit && transcript && seg_start && seg_len && nseg && overlap && cost && B > 0 && M > 0 && smax > 0,
                  "factk_match_cost: bad args");
    FACTK_REQUIRE(ldo >= M && M <= 4096, "factk_match_cost: bad token count %d / overlap stride %d", M, ldo);
    gt_overlap_kernel<<<dim3(smax, B), 128, 0, (cudaStream_t)stream>>>(attn, lda, aslot, ridx, seg_start, seg_len, nseg, overlap,
                                                                       smax, ldo, M, slot);
    int rc = check_launch("factk_match_cost(overlap)");
    if (rc) return rc;
    match_cost_kernel<<<B, 256, 2 * M * sizeof(float), (cudaStream_t)stream>>>(aclogit, M, C1, transcript, seg_len, nseg, overlap,
                                                                               smax, ldo, slot, pc, a2fc, cost, logp);
    return check_launch("factk_match_cost");
}

extern "C" int factk_token_loss(const float* aclogit, int M, int C1, const int32_t* aind, const int32_t* sind,
                                const int32_t* nmatch, int kmax, const int32_t* transcript, int smax, const float* cweight,
                                float* out, int out_stride, int B, int logp_mean, void* stream) {
    FACTK_REQUIRE(aclogit && aind && sind && nmatch && transcript && cweight && out && B > 0 && M > 0 && M <= 4096,
                  "factk_token_loss: bad args");
    token_loss_kernel<<<B, 256, 3 * M * sizeof(float), (cudaStream_t)stream>>>(aclogit, M, C1, aind, sind, nmatch, kmax, transcript,
                                                                               smax, cweight, out, out_stride, logp_mean);
    return check_launch("factk_token_loss");
}

extern "C" int factk_loss_combine(const float* ws, int nb, const int32_t* block_type, int B, int nchunk, const int32_t* len,
                                  const int32_t* npred, int C, int M, float sw, int use_clip, float fact_w, float con_w,
                                  int nseen, const int32_t* nvalid, float* out, int ldo, void* stream) {
    FACTK_REQUIRE(ws && block_type && len && out && B > 0 && nb > 0 && nb <= FACTK_LOSS_MAX_BLOCKS && ldo >= 4 + nb,
                  "factk_loss_combine: bad args (at most %d blocks)", FACTK_LOSS_MAX_BLOCKS);
    FACTK_REQUIRE(!use_clip || nvalid, "factk_loss_combine: the contrastive term needs the valid-frame counts");
    LossPlan plan;
    plan.nb = nb;
    for (int i = 0; i < nb; ++i) {
        FACTK_REQUIRE(block_type[i] >= 0 && block_type[i] <= 4 && (block_type[i] < 2 || npred), "factk_loss_combine: bad block type");
        plan.type[i] = block_type[i];
    }
    loss_combine_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ws, plan, B, nchunk, len, npred, C, M, sw, use_clip,
                                                                          fact_w, con_w, nseen, nvalid, out, ldo);
    return check_launch("factk_loss_combine");
}
