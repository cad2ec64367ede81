// Gradients of the training loss w.r.t. every logit tensor it reads (reference: models/loss.py:8-19, 196-341 and the
// compute_loss methods models/blocks.py:313-320, 369-382, 487-497, 677-786; the reference gets them from torch autograd).
// The loss VALUE kernels live in loss.cu; these are their mirrors.  Every kernel ADDS coef[b] * d(term)/dX into a
// zero-initialised gradient buffer of the logit tensor's layout; coef[b] carries the block mean, the FACT / InfoNCE mix and
// the 1/B of the batch mean.  One warp per row, fixed summation order, no atomics.
#include "common.cuh"

namespace factk {

__device__ __forceinline__ float warp_lse_row(const float* row, int K, int lane) {
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, row[k]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += expf(row[k] - mx);
    return mx + logf(warp_sum(s));
}

// frame_loss / frame_loss_tdu (loss.py:249-277): rows are frames (seg_start == NULL) or predicted segments whose target is
// the label histogram of their frames.  dX[r,c] += coef/Z * (p[r,c] * sum_t w[y_t]/len - sum_t w[y_t] [y_t == c]/len)
__global__ void ce_rows_grad_kernel(const float* __restrict__ X, int ldx, int C, float* __restrict__ dX, int lddx,
                                    const int32_t* __restrict__ label, const float* __restrict__ cw,
                                    const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_len,
                                    const int32_t* __restrict__ nrows, const float* __restrict__ coef, int slot) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * warps + (threadIdx.x >> 5), b = blockIdx.y;
    const int R = min(nrows[b], slot);
    if (r >= R) return;
    const size_t rb = (size_t)b * slot + r;
    const int t0 = seg_start ? seg_start[rb] : r, n = seg_start ? seg_len[rb] : 1;
    const float k = coef[b] / ((float)R * (float)n);
    const float* row = X + rb * ldx;
    const float lse = warp_lse_row(row, C, lane);
    float wsum = 0.f;
    for (int t = t0 + lane; t < t0 + n; t += 32) wsum += cw[label[(size_t)b * slot + t]];
    wsum = warp_sum(wsum);
    for (int c = lane; c < C; c += 32) {
        float hit = 0.f;
        for (int t = t0; t < t0 + n; ++t) hit += (label[(size_t)b * slot + t] == c) ? 1.f : 0.f;
        dX[rb * lddx + c] += k * (expf(row[c] - lse) * wsum - hit * cw[c]);
    }
}

// smooth_loss (loss.py:8-19) on logits: mean over (T-1) x C of clamp((lp[t+1,c] - lp[t,c])^2, 0, 16), lp = log_softmax.
__global__ void smooth_grad_kernel(const float* __restrict__ X, int ldx, int C, float* __restrict__ dX, int lddx,
                                   const int32_t* __restrict__ len, const float* __restrict__ coef, float mult, int slot) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * warps + (threadIdx.x >> 5), b = blockIdx.y;
    const int T = min(len[b], slot);
    if (t >= T || T < 2) return;
    const size_t rb = (size_t)b * slot + t;
    const float* row = X + rb * ldx;
    const float l1 = warp_lse_row(row, C, lane);
    const float l0 = t > 0 ? warp_lse_row(row - ldx, C, lane) : 0.f;
    const float l2 = t + 1 < T ? warp_lse_row(row + ldx, C, lane) : 0.f;
    const float k = 2.f * mult * coef[b] / ((float)(T - 1) * (float)C);
    float gsum = 0.f;
    // g[c] = dLoss/dlp[t,c]
    for (int c = lane; c < C; c += 32) {
        const float lp = row[c] - l1;
        float g = 0.f;
        if (t > 0) { const float d = lp - (row[c - ldx] - l0); if (d * d <= 16.f) g += d; }
        if (t + 1 < T) { const float d = (row[c + ldx] - l2) - lp; if (d * d <= 16.f) g -= d; }
        gsum += g;
    }
    gsum = warp_sum(gsum);
    for (int c = lane; c < C; c += 32) {
        const float lp = row[c] - l1;
        float g = 0.f;
        if (t > 0) { const float d = lp - (row[c - ldx] - l0); if (d * d <= 16.f) g += d; }
        if (t + 1 < T) { const float d = (row[c + ldx] - l2) - lp; if (d * d <= 16.f) g -= d; }
        dX[rb * lddx + c] += k * (g - expf(lp) * gsum);
    }
}

// action_token_loss (loss.py:196-209): class-weighted cross entropy, unmatched tokens labelled null (class C1-1), with
// one-to-many matches the last pair of a token wins.  dX[a,c] += coef * w[cl_a] / sum_a w[cl_a] * (p[a,c] - [c == cl_a])
__global__ void __launch_bounds__(256) token_ce_grad_kernel(const float* __restrict__ aclogit, int M, int C1, float* __restrict__ dX,
                                                            const int32_t* __restrict__ aind, const int32_t* __restrict__ sind,
                                                            const int32_t* __restrict__ nmatch, int kmax,
                                                            const int32_t* __restrict__ transcript, int smax,
                                                            const float* __restrict__ cweight, const float* __restrict__ coef) {
    extern __shared__ int clabel[];
    __shared__ float wtot;
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int a = threadIdx.x; a < M; a += blockDim.x) clabel[a] = C1 - 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < nmatch[b]; ++k)
            clabel[aind[(size_t)b * kmax + k]] = transcript[(size_t)b * smax + sind[(size_t)b * kmax + k]];
        float d = 0.f;
        for (int a = 0; a < M; ++a) d += cweight[clabel[a]];
        wtot = d;
    }
    __syncthreads();
    for (int a = wid; a < M; a += 8) {
        const float* row = aclogit + ((size_t)b * M + a) * C1;
        const float lse = warp_lse_row(row, C1, lane);
        const float k = coef[b] * cweight[clabel[a]] / wtot;
        for (int c = lane; c < C1; c += 32)
            dX[((size_t)b * M + a) * C1 + c] += k * (expf(row[c] - lse) - (c == clabel[a] ? 1.f : 0.f));
    }
}

// cross_attn_loss / cross_attn_loss_tdu (loss.py:211-247).  Written as a sum over frames t (like loss_pick):
//   loss = 1/Z * sum_t w[gseg(t)] / rlen(t) * (lse(t) - X[r(t), col[gseg(t)]])
// with col[gs] = the token matched to ground-truth segment gs (-1: unmatched), w[gs] the segment weight at its match
// position, r(t) = the frame itself or its predicted segment, Z = T or the predicted segment count.
//   mode 0 (a2f, softmax over the matched tokens of a row; mult[m] = how often token m occurs among the matches):
//       dX[r,m] += G_r * mult[m] exp(X[r,m]) / sum_m' mult[m'] exp(X[r,m']) - q[r,m]
//   mode 1 (f2a, softmax over the rows of a column; colmass[m] = sum of the frame weights whose target column is m):
//       dX[r,m] += colmass[m] * exp(X[r,m] - lse_col[m]) - q[r,m]
// q[r,m] = sum over the row's frames whose target column is m of their weight, G_r = sum_m q[r,m]: both from the overlap
// of the row's frame range with the (contiguous) ground-truth segments.
__global__ void xattn_grad_kernel(int mode, const float* __restrict__ X, int ldx, int M, float* __restrict__ dX, int lddx,
                                  const int32_t* __restrict__ gseg, const int32_t* __restrict__ gstart, const int32_t* __restrict__ glen,
                                  const int32_t* __restrict__ colmap, const float* __restrict__ wmap, int smax,
                                  const float* __restrict__ mult, const float* __restrict__ colmass, const float* __restrict__ col_lse,
                                  int ld_lse, const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_len,
                                  const int32_t* __restrict__ nrows, const float* __restrict__ coef, int slot) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * warps + (threadIdx.x >> 5), b = blockIdx.y;
    const int R = min(nrows[b], slot);
    if (r >= R) return;
    const size_t rb = (size_t)b * slot + r;
    const int fs = seg_start ? seg_start[rb] : r, n = seg_start ? seg_len[rb] : 1, fe = fs + n;
    const float k = coef[b] / ((float)R * (float)n);
    const int g0 = gseg[(size_t)b * slot + fs], g1 = gseg[(size_t)b * slot + fe - 1];
    const int32_t* cm = colmap + (size_t)b * smax;
    const float* wm = wmap + (size_t)b * smax;
    const float* row = X + rb * ldx;
    float G = 0.f;
    if (mode == 0) {
        for (int gs = g0 + lane; gs <= g1; gs += 32) {
            if (cm[gs] < 0) continue;
            const int s0 = gstart[(size_t)b * slot + gs], s1 = s0 + glen[(size_t)b * slot + gs];
            G += (float)(min(fe, s1) - max(fs, s0)) * wm[gs];
        }
        G = warp_sum(G) * k;
    }
    float mx = -INFINITY, zs = 0.f;
    if (mode == 0) {
        const float* mu = mult + (size_t)b * M;
        for (int m = lane; m < M; m += 32) if (mu[m] > 0.f) mx = fmaxf(mx, row[m]);
        mx = warp_max(mx);
        for (int m = lane; m < M; m += 32) if (mu[m] > 0.f) zs += mu[m] * expf(row[m] - mx);
        zs = warp_sum(zs);
    }
    // the usual row lies inside one ground-truth segment or straddles two: their (column, weight) pairs are fetched once, not per column
    const bool few = g1 - g0 <= 1;
    int c0 = -2, c1 = -2;
    float q0 = 0.f, q1 = 0.f;
    if (few) {
        c0 = cm[g0];
        const int a0 = gstart[(size_t)b * slot + g0], a1 = a0 + glen[(size_t)b * slot + g0];
        q0 = (float)(min(fe, a1) - max(fs, a0)) * wm[g0];
        if (g1 > g0) {
            c1 = cm[g1];
            const int b0 = gstart[(size_t)b * slot + g1], b1 = b0 + glen[(size_t)b * slot + g1];
            q1 = (float)(min(fe, b1) - max(fs, b0)) * wm[g1];
        }
    }
#pragma unroll 2
    for (int m = lane; m < M; m += 32) {
        float q = 0.f;
        if (few) {
            q = (m == c0 ? q0 : 0.f) + (m == c1 ? q1 : 0.f);
        } else {
            for (int gs = g0; gs <= g1; ++gs) {
                if (cm[gs] != m) continue;
                const int s0 = gstart[(size_t)b * slot + gs], s1 = s0 + glen[(size_t)b * slot + gs];
                q += (float)(min(fe, s1) - max(fs, s0)) * wm[gs];
            }
        }
        q *= k;
        float g;
        if (mode == 0) {
            const float mu = mult[(size_t)b * M + m];
            g = (mu > 0.f ? G * mu * expf(row[m] - mx) / zs : 0.f) - q;
        } else {
            const float cmass = colmass[(size_t)b * M + m];
            g = (cmass != 0.f ? cmass * expf(row[m] - col_lse[(size_t)b * ld_lse + m]) : 0.f) - q;
        }
        if (g != 0.f) dX[rb * lddx + m] += g;
    }
}

// colmass[b][m] = coef/Z * sum over ground-truth segments gs with col[gs] == m of w[gs] * sum_{t in gs} 1/rlen(t)
// One CTA per (column, video): the frames of a matching segment are spread over the threads (a column that owns most of a long
// video would otherwise be one thread walking 16384 frames); fixed thread -> frame assignment and block_sum: deterministic.
__global__ void __launch_bounds__(128) xattn_colmass_kernel(float* __restrict__ colmass, int M, const int32_t* __restrict__ gstart,
                                                            const int32_t* __restrict__ glen, const int32_t* __restrict__ gn,
                                                            const int32_t* __restrict__ colmap, const float* __restrict__ wmap, int smax,
                                                            const int32_t* __restrict__ seg_label, const int32_t* __restrict__ seg_len,
                                                            const int32_t* __restrict__ nrows, const float* __restrict__ coef, int slot) {
    __shared__ float sm[32];
    const int m = blockIdx.x, b = blockIdx.y;
    const float k = coef[b] / (float)min(nrows[b], slot);
    float s = 0.f;
    const int n_gs = gn[b];
    for (int gs = 0; gs < n_gs; ++gs) {
        if (colmap[(size_t)b * smax + gs] != m) continue;
        const int s0 = gstart[(size_t)b * slot + gs], n = glen[(size_t)b * slot + gs];
        const float w = wmap[(size_t)b * smax + gs];
        if (seg_label) {
            for (int t = s0 + threadIdx.x; t < s0 + n; t += blockDim.x)
                s += w / (float)seg_len[(size_t)b * slot + seg_label[(size_t)b * slot + t]];
        } else if (threadIdx.x == 0) {
            s += w * (float)n;
        }
    }
    s = block_sum(s, sm);
    if (threadIdx.x == 0) colmass[(size_t)b * M + m] = s * k;
}

// infonce_contrastive_loss (loss.py:280-341) on sim = emb . text^T / temp, restricted to the seen classes (cmap[c] >= 0) and
// to the frames with a seen label (blocks.py:697-748):  loss = (CE_rows + mean_c CE_cols(c) / max(count_c, 1)) / 2
__global__ void class_count_kernel(const int32_t* __restrict__ label, const int32_t* __restrict__ len, int C, float* __restrict__ count,
                                   int slot) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= C) return;
    int n = 0;
    for (int t = 0; t < min(len[b], slot); ++t) n += label[(size_t)b * slot + t] == c;
    count[(size_t)b * C + c] = (float)n;
}

__global__ void infonce_grad_kernel(const float* __restrict__ sim, int lds, int C, float* __restrict__ dS, int ldds,
                                    const int32_t* __restrict__ label, const int32_t* __restrict__ cmap, const float* __restrict__ count,
                                    const float* __restrict__ col_lse, int ld_lse, const int32_t* __restrict__ nvalid, int nseen,
                                    const int32_t* __restrict__ len, const float* __restrict__ coef, int slot) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * warps + (threadIdx.x >> 5), b = blockIdx.y;
    if (t >= min(len[b], slot)) return;
    const size_t rb = (size_t)b * slot + t;
    const int y = label[rb];
    if (cmap[y] < 0 || nvalid[b] <= 0) return;               // frames of held-out classes leave the loss
    const float* row = sim + rb * lds;
    float mx = -INFINITY, zs = 0.f;
    for (int c = lane; c < C; c += 32) if (cmap[c] >= 0) mx = fmaxf(mx, row[c]);
    mx = warp_max(mx);
    for (int c = lane; c < C; c += 32) if (cmap[c] >= 0) zs += expf(row[c] - mx);
    zs = warp_sum(zs);
    const float k1 = 0.5f * coef[b] / (float)nvalid[b], k2 = 0.5f * coef[b] / (float)nseen;
    for (int c = lane; c < C; c += 32) {
        if (cmap[c] < 0) continue;
        const float cnt = count[(size_t)b * C + c];
        float g = k1 * (expf(row[c] - mx) / zs - (c == y ? 1.f : 0.f));
        if (cnt > 0.f) g += k2 * (expf(row[c] - col_lse[(size_t)b * ld_lse + c]) - (c == y ? 1.f / cnt : 0.f));
        dS[rb * ldds + c] += g;
    }
}

}  // namespace factk

using namespace factk;

/* frame_loss (seg_start == NULL: rows are frames, nrows = len) or frame_loss_tdu (rows = predicted segments, nrows = nseg). */
extern "C" int factk_loss_grad_ce_rows(const float* X, int ldx, int C, float* dX, int lddx, const int32_t* label, const float* cweight,
                                       const int32_t* seg_start, const int32_t* seg_len, const int32_t* nrows, const float* coef,
                                       int B, int slot, void* stream) {
    FACTK_REQUIRE(X && dX && label && cweight && nrows && coef && C > 0 && (!seg_start || seg_len), "factk_loss_grad_ce_rows: bad args");
    ce_rows_grad_kernel<<<dim3((slot + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(X, ldx, C, dX, lddx, label, cweight, seg_start, seg_len,
                                                                                 nrows, coef, slot);
    return check_launch("factk_loss_grad_ce_rows");
}

extern "C" int factk_loss_grad_smooth(const float* X, int ldx, int C, float* dX, int lddx, const int32_t* len, const float* coef,
                                      float mult, int B, int slot, void* stream) {
    FACTK_REQUIRE(X && dX && len && coef && C > 0, "factk_loss_grad_smooth: bad args");
    smooth_grad_kernel<<<dim3((slot + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(X, ldx, C, dX, lddx, len, coef, mult, slot);
    return check_launch("factk_loss_grad_smooth");
}

extern "C" int factk_loss_grad_token(const float* aclogit, int M, int C1, float* dX, const int32_t* aind, const int32_t* sind,
                                     const int32_t* nmatch, int kmax, const int32_t* transcript, int smax, const float* cweight,
                                     const float* coef, int B, void* stream) {
    FACTK_REQUIRE(aclogit && dX && aind && sind && nmatch && transcript && cweight && coef && M > 0 && M <= 4096,
                  "factk_loss_grad_token: bad args");
    token_ce_grad_kernel<<<B, 256, M * sizeof(int), (cudaStream_t)stream>>>(aclogit, M, C1, dX, aind, sind, nmatch, kmax, transcript, smax,
                                                                            cweight, coef);
    return check_launch("factk_loss_grad_token");
}

/* mode 0: a2f direction (softmax over the matched tokens per row; mult required); mode 1: f2a direction (softmax over rows per
 * column; col_lse required, colmass = scratch [B][M]).  seg_* == NULL: rows are frames (nrows = len); else predicted segments. */
extern "C" int factk_loss_grad_xattn(int mode, const float* X, int ldx, int M, float* dX, int lddx, const int32_t* gseg,
                                     const int32_t* gstart, const int32_t* glen, const int32_t* gn, const int32_t* colmap,
                                     const float* wmap, int smax, const float* mult, float* colmass, const float* col_lse, int ld_lse,
                                     const int32_t* seg_label, const int32_t* seg_start, const int32_t* seg_len, const int32_t* nrows,
                                     const float* coef, int B, int slot, void* stream) {
    FACTK_REQUIRE(X && dX && gseg && gstart && glen && gn && colmap && wmap && nrows && coef && M > 0, "factk_loss_grad_xattn: bad args");
    FACTK_REQUIRE(mode == 0 ? mult != nullptr : (colmass && col_lse), "factk_loss_grad_xattn: mode %d operands missing", mode);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1)
        xattn_colmass_kernel<<<dim3(M, B), 128, 0, st>>>(colmass, M, gstart, glen, gn, colmap, wmap, smax, seg_label, seg_len,
                                                                   nrows, coef, slot);
    xattn_grad_kernel<<<dim3((slot + 7) / 8, B), 256, 0, st>>>(mode, X, ldx, M, dX, lddx, gseg, gstart, glen, colmap, wmap, smax, mult,
                                                             colmass, col_lse, ld_lse, seg_start, seg_len, nrows, coef, slot);
    return check_launch("factk_loss_grad_xattn");
}

/* count: scratch [B][C]; col_lse [B][ld_lse]: log-sum-exp of every class column over the frames with a seen label (factk_col_lse). */
extern "C" int factk_loss_grad_infonce(const float* sim, int lds, int C, float* dS, int ldds, const int32_t* label, const int32_t* cmap,
                                       float* count, const float* col_lse, int ld_lse, const int32_t* nvalid, int nseen,
                                       const int32_t* len, const float* coef, int B, int slot, void* stream) {
    FACTK_REQUIRE(sim && dS && label && cmap && count && col_lse && nvalid && len && coef && C > 0 && nseen > 0,
                  "factk_loss_grad_infonce: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    class_count_kernel<<<dim3((C + 63) / 64, B), 64, 0, st>>>(label, len, C, count, slot);
    infonce_grad_kernel<<<dim3((slot + 7) / 8, B), 256, 0, st>>>(sim, lds, C, dS, ldds, label, cmap, count, col_lse, ld_lse, nvalid, nseen,
                                                               len, coef, slot);
    return check_launch("factk_loss_grad_infonce");
}
