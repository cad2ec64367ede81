/*
 * factk -- C ABI of the B200 (sm_100a) kernel library behind the FACT / FACT_CLIP forward.
 *
 * The reference (lucas-t-t/FACT-CLIP) has no FFI / plugin layer: its boundary is the Python
 * nn.Module API of fact_clip/models/blocks.py (FACT :19-135, FACT_CLIP :504-920).  The host side in
 * fact_clip_b200/ mirrors that API and binds THIS header through ctypes (INTEGRATION.md shows the
 * stub).  Every entry point below names the reference code it replaces (path:line under
 * /root/reference/fact_clip/).
 *
 * Conventions
 *  - all pointers are BORROWED device pointers (except factk_last_error); the library never
 *    allocates or frees caller memory and never synchronises; every launch goes to `stream`
 *    (a cudaStream_t passed as void*).
 *  - return value: 0 on success, negative on error; factk_last_error() gives a thread-local text.
 *  - "rows" layout: an activation is a row-major matrix [B][slot][ld] -- B videos, `slot` row
 *    slots per video of which the first len[b] are valid (len == NULL: all slot rows valid), `ld`
 *    elements per row.  Frames, segments and action tokens all use it.  Rows >= len[b] are never
 *    read as data (they read as zero for convolution taps) and never written.
 *  - dtype codes: FACTK_F32 = 0, FACTK_BF16 = 1.
 */
#ifndef FACTK_H_
#define FACTK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FACTK_F32 0
#define FACTK_BF16 1
#define FACTK_MAX_SRC 6
#define FACTK_LOSS_MAX_BLOCKS 8

#define FACTK_OK 0
#define FACTK_ERR_ARG (-1)
#define FACTK_ERR_CUDA (-2)
#define FACTK_ERR_DEVICE (-3)
#define FACTK_ERR_UNSUPPORTED (-4)

int factk_version(void);
const char* factk_last_error(void);
/* 0 when the current device is sm_100 (B200); FACTK_ERR_DEVICE otherwise. */
int factk_device_check(void);

/* One operand of the generalised GEMM: rows of A (optionally shifted / gathered / with a positional
 * term on the first pos_d channels) times W^T.  Replaces, depending on use: nn.Conv1d taps
 * (models/basic.py:138-139,158,177,182), nn.Linear (basic.py:343-347, blocks.py:402,414,153-159) and
 * add_positional_encoding (basic.py:313-320). */
typedef struct {
    const void* A;        /* [B][a_slot][lda] */
    int32_t a_dtype;
    int32_t lda;
    int32_t a_slot;
    int32_t K;
    int32_t row_off;      /* tap offset in rows; source rows outside [0,len[b]) read as zero */
    int32_t pos_ld;
    int32_t pos_d;
    int32_t ldw;
    int32_t w_dtype;      /* FACTK_F32 for factk_gemm; FACTK_BF16 (or FACTK_F32 = tf32 math) for factk_gemm_tc */
    int32_t reserved_;
    const int32_t* gather;  /* optional [B][slot] source-row index (seg->frame upsample, blocks.py:442) */
    const float* pos;       /* optional table [n][pos_ld]; added to channels < pos_d */
    const int32_t* pos_idx; /* optional [B][slot] row index into pos (seg_center, blocks.py:454-455); default: row */
    const void* W;          /* [N][ldw] of w_dtype, y = A W^T */
    int64_t w_bstride;      /* elements between per-video weights; 0 = shared */
} factk_src_t;

typedef struct {
    int32_t B, slot, N, nsrc;
    const int32_t* len;     /* [B] valid rows, or NULL */
    factk_src_t src[FACTK_MAX_SRC];
    const float* bias;      /* [N] or NULL */
    int64_t bias_bstride;
    float alpha;            /* y = relu?(alpha * sum + bias + pre) + res */
    int32_t relu;
    const void* pre;        /* optional pre-activation addend rows: pre[b*pre_bstride + idx(t)*ldpre + n] */
    const int32_t* pre_idx; /* optional [B][slot] row index into pre (default: t) */
    int64_t pre_bstride;    /* elements between per-video blocks of pre; 0 = shared table */
    int32_t pre_dtype, ldpre;
    const void* res;        /* optional residual [B][slot][ldres], added after the activation */
    int32_t res_dtype, ldres;
    void* Y;                /* [B][slot][ldy] */
    int32_t y_dtype, ldy;
} factk_gemm_t;

/* y[b,t,:N] = act(alpha * sum_s A_s[b, idx_s(t)] W_s[b]^T + bias + pre) + res.  fp32 accumulate (CUDA cores). */
int factk_gemm(const factk_gemm_t* g, void* stream);

/* Same contract on the 5th-generation tensor cores: TMA-staged operands (128B-swizzled shared memory),
 * tcgen05.mma with fp32 accumulators in tensor memory, epilogue from TMEM.  Restrictions (else
 * FACTK_ERR_UNSUPPORTED): every source has the same dtype -- bf16 A with bf16 W (kind::f16), or fp32 A
 * with fp32 W (kind::tf32, operands truncated to tf32 by the tensor core); K a multiple of 64 (bf16) /
 * 32 (fp32); no gather / pos on sources (use `pre`); 16-byte aligned rows; slot a multiple of 128.
 * Rows of A in [len[b], slot) must be zero where convolution taps (row_off != 0) can reach them. */
int factk_gemm_tc(const factk_gemm_t* g, void* stream);
/* 1 if factk_gemm_tc accepts this descriptor, 0 otherwise (no launch). */
int factk_gemm_tc_supported(const factk_gemm_t* g);

/* Frame-side bf16 GEMM on CTA pairs (tcgen05 cta_group::2, two TMEM accumulators, bf16-staged coalesced epilogue):
 *   y[b,t,:N] = act(x[b,t,:K] W^T + bias + pre[b*pre_bstride + idx(t)*ldpre + :N]),  x / W / y bf16, bias / pre fp32.
 * Replaces the plain Linear / 1x1 Conv1d call sites on frames and segments: conv_out (basic.py:182,213), the SCA key /
 * value projection (basic.py:465,513), sf_merge (blocks.py:414,445, `pre` = the gathered segment term), the CLIP projection
 * (blocks.py:153-159), seg_combine (blocks.py:402).  K % 64 == 0, N % 128 == 0 (factk_gemm_pair_supported), 16-byte aligned
 * rows; rows >= len[b] are not written. */
int factk_gemm_pair_supported(int K, int N);
int factk_gemm_pair(const void* x, int lda, int a_slot, const void* W, int ldw, int K, int N, const float* bias,
                    const float* pre, int ldpre, long long pre_bstride, const int32_t* pre_idx, int relu,
                    void* y, int ldy, int B, int slot, const int32_t* len, void* stream);

/* Fused DilatedResidualLayer (models/basic.py:154-171; MSTCN.forward :216-217), eval mode:
 *   y[b,t,:] = x[b,t,:] + W1 relu(sum_k W3[k] x[b, t+(k-1)*dilation, :] + b3) + b1
 * in ONE persistent tcgen05 kernel (conv3 -> ReLU tile kept in shared memory -> 1x1 -> residual), nothing but x
 * in and y out touches HBM.  x, y: bf16 [B][slot][F] contiguous, y != x, rows of x in [len[b], slot) must be zero
 * (convolution zero padding; rows of y >= len[b] are not written); w3: bf16 [3][F][F] = Conv1d weight
 * (F,F,3) permuted to (tap, out, in); w1: bf16 [F][F]; b3, b1: fp32 [F].  F in {128, 256}
 * (factk_tcn_layer_supported).  cta_group: 2 = tcgen05 cta_group::2 CTA pairs (default), 1 = single CTA. */
int factk_tcn_layer_supported(int F);
int factk_tcn_layer(const void* x, void* y, const void* w3, const float* b3, const void* w1, const float* b1,
                    int B, int slot, int F, int dilation, const int32_t* len, int cta_group, void* stream);
/* Same, additionally recording a clock64 timeline of CTA 0 into dbg[64][16] (development aid; dbg may be NULL). */
int factk_tcn_layer_dbg(const void* x, void* y, const void* w3, const float* b3, const void* w1, const float* b1,
                        int B, int slot, int F, int dilation, const int32_t* len, int cta_group, long long* dbg, void* stream);

/* Block.process_feature (blocks.py:195-202) in place on rows [B][slot][ld] of width H: the last C
 * channels are logits -> clogit_out fp32 [B][slot][C]; they are overwritten by softmax(logits).
 * pred_out (optional, int32 [B][slot]) = argmax of the probabilities, first index on ties
 * (blocks.py:420-421). */
int factk_softmax_splice(void* X, int dtype, int B, int slot, const int32_t* len, int ld, int H, int C,
                         float* clogit_out, int32_t* pred_out, void* stream);

/* y = LayerNorm(x (+ r)) * w + b, optional ReLU (nn.LayerNorm in basic.py:407-408,472-474,
 * blocks.py:155,223).  x, r: [B][slot][ld*] of dtype; y may alias x. */
int factk_layernorm(const void* X, int x_dtype, int ldx, const void* R, int r_dtype, int ldr,
                    const float* w, const float* b, float eps, int relu,
                    void* Y, int y_dtype, int ldy, int B, int slot, const int32_t* len, int E, void* stream);

/* F.normalize(x, dim=-1, eps) (blocks.py:174). */
int factk_l2norm(const void* X, int x_dtype, int ldx, void* Y, int y_dtype, int ldy,
                 int B, int slot, const int32_t* len, int E, float eps, void* stream);

/* softmax over the first M columns of each row: P = softmax(scale * L) (basic.py:376, a2f direction;
 * blocks.py:826 with scale = 1).  L fp32 [B][slot][ldl]; P fp32 [B][slot][ldp] (may alias L).
 * P16 (optional): bf16 copy [B][slot][ldp16] with columns M..pad16-1 zeroed -- the tensor-core operand. */
int factk_row_softmax(const float* L, int ldl, float* P, int ldp, int B, int slot, const int32_t* len, int M,
                      float scale, void* P16, int ldp16, int pad16, void* stream);

/* Token self-attention core of nn.MultiheadAttention (basic.py:500, 437): Q,K,V fp32 [B][M][ld]
 * already projected; O[b,m,h*dh:(h+1)*dh] = softmax(q k^T / sqrt(dh)) v per head.  tf32 != 0: tensor-core kernel
 * (tf32 operands, fp32 accumulate; dh 16/32/64), the bf16 compute mode's choice; 0: fp32 CUDA-core kernel. */
int factk_mha_tokens(const float* Q, const float* K, const float* V, int ld, float* O, int ldo,
                     int B, int M, int nhead, int dh, int tf32, void* stream);

/* Tokens-attend-rows core of the SCALayer cross attention (basic.py:507-514): Q fp32 [B][M][ldq]
 * (projected), Kx/Vx rows [B][slot][ldkv] of dtype (projected keys / values, column offsets applied
 * by the caller), softmax over the len[b] valid rows per head.  ws: fp32 scratch of
 * factk_attn_rows_ws_floats(...) floats. */
size_t factk_attn_rows_ws_floats(int B, int slot, int M, int nhead, int dh);
int factk_attn_rows(const float* Q, int ldq, const void* Kx, const void* Vx, int kv_dtype, int ldkv,
                    float* O, int ldo, int B, int slot, const int32_t* len, int M, int nhead, int dh,
                    float* ws, void* stream);

/* Development aid of the tcgen05 attention kernel behind factk_attn_rows (csrc/attn_tc.cu): clock64 timeline of one CTA. */
int factk_attn_tc_debug(long long* dbg);

/* Softmax over ROWS (the X axis of X2Y_map in the f2a direction, basic.py:373-379):
 *   L fp32 [B][slot][ldl] holds logits (row = frame/segment, column = token m < M);
 *   stats[b][m] = (max, sum exp) over valid rows;
 *   out[b][m][:E] = sum_t softmax_t(L[b,t,m]) * X[b,t,:E]   (X of dtype, row-major [B][slot][ldx]);
 *   P (optional, fp32 [B][slot][ldp]) receives the normalised attention. */
size_t factk_col_softmax_ws_floats(int B, int slot, int M, int E);
int factk_col_softmax_apply(const float* L, int ldl, const void* X, int x_dtype, int ldx,
                            float* out, int ldo, float* P, int ldp,
                            int B, int slot, const int32_t* len, int M, int E, float* ws, void* stream);

/* X2Y_map in the a2f direction (basic.py:349-389 with X = tokens, Y = rows) as ONE tcgen05 / TMEM kernel per 128-row tile
 * (csrc/x2y_fused.cu): S = rows . kt^T + cb -> softmax over the tokens with the row of S in registers -> P (bf16, shared
 * memory) -> out = rows . Wy^T + P . vt^T + bias; the rows are read once, logits / attention reach HBM only when the
 * pointers are given (loss, eval fusion).  rows X bf16 [B][slot][ldx] (K = H), kt bf16 [B][M][ldkt] (the token-side fold
 * alpha Wq^T X_K(tokens)), cb fp32 [B][M], Wy bf16 [F][ldwy], vt bf16 [B][F][ldvt] (columns >= M zero up to a multiple of 64),
 * bias fp32 [F], out bf16 [B][slot][ldo], logit / attn fp32 [B][slot][ldl] or NULL.  H % 64 == 0, F == 256, M <= 128. */
int factk_a2f_fused_supported(int M, int H, int F, int slot);
int factk_a2f_fused(const void* X, int ldx, const void* Kt, int ldkt, long long kt_bstride, const float* cb, long long cb_bstride,
                    const void* Wy, int ldwy, const void* Vt, int ldvt, long long vt_bstride, const float* bias, void* out, int ldo,
                    float* logit, float* attn, int ldl, int B, int slot, const int32_t* len, int M, int H, int F, void* stream);

/* X2Y_map, f2a direction (models/basic.py:349-389 with X = frame / segment rows, Y = tokens; blocks.py:346,458) as ONE
 * tcgen05 / TMEM kernel + a fixed-order combine: out[b][m][:] = sum_t softmax_t(qt[b][m] . X[b][t]) X[b][t][:], t < len[b].
 * The rows are read once; the [B][slot][M] logits never reach HBM (the per-token logit bias is constant along t and cancels).
 * rows X bf16 [B][slot][ldx] (H channels), qt bf16 [B][M][ldq] (the token-side fold alpha Wk^T Y_Q(tokens)), out fp32 [B][M][ldo],
 * ws >= factk_f2a_fused_ws_floats(B, slot, M, H) floats.  H == 512 with M <= 80 or H == 256 with M <= 128; slot % 128 == 0.
 * Rows in [len[b], slot) may hold anything (masked; zeroed in shared memory before the value product of a ragged tile).
 * Kernels: f2a_fused_t.cu (rows on the M side of both MMAs; H = 512, and H = 256 up to 80 tokens), f2a_fused.cu (tokens on the
 * M side; H = 256).  Results are bit-reproducible and do not depend on the batch a video is in. */
int factk_f2a_fused_supported(int M, int H, int slot);
size_t factk_f2a_fused_ws_floats(int B, int slot, int M, int H);
int factk_f2a_fused(const void* X, int ldx, const void* Qt, int ldq, long long qt_bstride, float* out, int ldo, int B, int slot,
                    const int32_t* len, int M, int H, float* ws, void* stream);
/* development aid: clock64 timeline of one CTA of the next factk_f2a_fused launches (tools/bench_f2a.py); NULL = off */
int factk_f2a_debug(long long* dbg);

/* Run-length segmentation on device (utils/utils.py:25-48, basic.py:597-607, blocks.py:454):
 * pred int32 [B][slot] -> seg_label[B][slot], seg_start[B][slot], seg_len[B][slot],
 * seg_center[B][slot] (= (start+end)/2 floor), nseg[B]. */
int factk_tdu_segment(const int32_t* pred, int B, int slot, const int32_t* len,
                      int32_t* seg_label, int32_t* seg_start, int32_t* seg_len, int32_t* seg_center,
                      int32_t* nseg, void* stream);

/* TemporalDownsampleUpsample.feature_frame2seg (basic.py:615-625): deterministic segment mean.
 * X rows of dtype [B][slot][ldx] -> seg fp32/bf16 [B][slot][lds], first nseg[b] rows.  seg_label / seg_start / seg_len /
 * nseg as written by factk_tdu_segment.  ws: fp32 scratch of factk_segment_mean_ws_floats(...) floats for the streaming
 * two-pass kernel (every frame read once, time independent of the segmentation); ws == NULL selects the one-pass kernel. */
size_t factk_segment_mean_ws_floats(int B, int slot, int E);
int factk_segment_mean(const void* X, int x_dtype, int ldx, void* seg, int s_dtype, int lds,
                       const int32_t* seg_label, const int32_t* seg_start, const int32_t* seg_len, const int32_t* nseg,
                       int B, int slot, int E, float* ws, void* stream);

/* Bidirectional GRU recurrence (nn.GRU(H, H/2, 1, bidirectional=True), blocks.py:401,432) over the
 * nseg[b] segments of each video.  gi fp32 [B][slot][6*Hh] = x W_ih^T + b_ih for (fwd r,z,n | bwd r,z,n);
 * w_hh_{f,b} fp32 [3*Hh][Hh]; b_hh_{f,b} [3*Hh].  out [B][slot][2*Hh] of dtype = relu?(cat[fwd,bwd]). */
int factk_gru_bidir(const float* gi, const float* w_hh_f, const float* b_hh_f,
                    const float* w_hh_b, const float* b_hh_b, int Hh,
                    void* out, int o_dtype, int ldo, int relu,
                    int B, int slot, const int32_t* nseg, void* stream);

/* Same recurrence on the tensor cores for the bf16 compute mode: 8 videos of a direction advance together as one
 * [768 x 256] x [256 x 8] mma.sync product per step on an 8-CTA cluster (W_hh as bf16 A fragments in registers, hidden
 * state exchanged as bf16 through distributed shared memory, fp32 state / gates with MUFU.TANH).  Hh must be 256. */
int factk_gru_bidir_mma(const float* gi, const float* w_hh_f, const float* b_hh_f,
                        const float* w_hh_b, const float* b_hh_b, int Hh,
                        void* out, int o_dtype, int ldo, int relu,
                        int B, int slot, const int32_t* nseg, void* stream);
/* Same, recording a clock64 timeline of CTA 0 into dbg[64][16] (development aid; dbg may be NULL). */
/* The same with the videos assigned to the 8-video chain groups in order of decreasing segment count (ranked on the
 * device into order_ws, int32 [B]): chains of similar length share a cluster, whole clusters retire early, and the
 * remaining ones step faster (2.32 -> 2.08 ms at 64 videos with 564..2572 segments).  Results are identical. */
int factk_gru_bidir_mma_sorted(const float* gi, const float* w_hh_f, const float* b_hh_f,
                               const float* w_hh_b, const float* b_hh_b, int Hh, void* out, int o_dtype, int ldo,
                               int relu, int B, int slot, const int32_t* nseg, int32_t* order_ws, void* stream);

/* Diagnostic: number of 8-CTA clusters of the tensor-core GRU kernel that can be co-resident on the current device. */
int factk_gru_max_clusters(void);

int factk_gru_bidir_mma_dbg(const float* gi, const float* w_hh_f, const float* b_hh_f,
                            const float* w_hh_b, const float* b_hh_b, int Hh,
                            void* out, int o_dtype, int ldo, int relu,
                            int B, int slot, const int32_t* nseg, long long* dbg, void* stream);

/* out[b,t,:E] = in[b, idx[b,t], :E] fp32 (attn_seg2frame, basic.py:645-651). */
int factk_gather_rows(const float* in, int ldi, int in_slot, const int32_t* idx, float* out, int ldo,
                      int B, int slot, const int32_t* len, int E, void* stream);

/* Prob fusion + argmax: Block._eval (blocks.py:242-261) / FACT_CLIP.eval_with_clip (blocks.py:854-887).
 * action_clogit fp32 [B][M][C+1]; attn fp32 [B][attn_slot][lda] (a2f attention, row = frame, or row =
 * segment when seg_label != NULL); flogit fp32 [B][slot][ldf] frame-branch logits (frame_clogit, or
 * CLIP similarity / temp) -> pred int64 [B][slot].  f_logp != 0: flogit rows are log-probabilities used as exp(.)
 * WITHOUT re-normalisation (blocks_SepVerbNoun.py:307-329, where they are verb x noun products). */
int factk_fuse_eval(const float* action_clogit, const float* attn, int lda, int attn_slot,
                    const int32_t* seg_label, const float* flogit, int ldf, float weight,
                    int64_t* pred, int B, int slot, const int32_t* len, int M, int C, int f_logp, void* stream);

/* Block._eval_w_transcript (blocks.py:263-275), FACT.trans models (the tokens are the video's transcript): per frame
 * prob[n] = (1 - weight) softmax_n(attn[row, :N]) + weight softmax_c(flogit)[transcript[n]], pred = transcript[argmax_n].
 * attn fp32 [B][attn_slot][lda] (row = frame, or segment when seg_label != NULL); transcript int32 [B][ldt], ntr[b] entries. */
int factk_fuse_eval_transcript(const float* attn, int lda, int attn_slot, const int32_t* seg_label,
                               const float* flogit, int ldf, float weight,
                               const int32_t* transcript, int ldt, const int32_t* ntr,
                               int64_t* pred, int B, int slot, const int32_t* len, int C, void* stream);

/* Token initialisation of FACT.trans models (blocks.py:74-79): out[n,:A] = embed[transcript[n],:A] + pe[n,:A]. */
int factk_embed_tokens(const float* embed, int lde, const int32_t* transcript, const float* pe, int ldpe,
                       float* out, int ldo, int N, int A, void* stream);

/* Input staging: dst[b][t][d] = src_b[d][t] for t < len[b], src_b = src + b * src_bstride a dense [D][len[b]] fp32 array
 * (features stored channel-major, which the reference transposes on the host: utils/dataset.py:12-21); dst fp32 or bf16. */
int factk_transpose_rows(const float* src, long long src_bstride, void* dst, int dst_dtype, int ldd, int B, int slot, int D,
                         const int32_t* len, void* stream);

/* Epic verb/noun heads (blocks_SepVerbNoun.py).  factk_vn_splice: process_feature (:229-234) in place on the last n1+n2
 * channels of X (two softmaxes), raw logits to clogit_out [B][slot][n1+n2]; with pred_out, also the segmentation argmax
 * over the nact action classes a = (vids[a], nids[a]) of verb-prob x noun-prob (:281-290). */
int factk_vn_splice(void* X, int dtype, int B, int slot, const int32_t* len, int ld, int H, int n1, int n2,
                    float* clogit_out, const int32_t* vids, const int32_t* nids, int nact, int32_t* pred_out, void* stream);

/* combine_verb_noun_to_action(apply_log=True) (:188-226): out[b][t][a] = log_softmax(clogit[:k1])[vids[a]] +
 * log_softmax(clogit[k1:k1+k2])[nids[a]]; with_null appends the null action (last class of both heads). */
int factk_vn_combine(const float* clogit, int ldc, int k1, int k2, const int32_t* vids, const int32_t* nids, int nact,
                     int with_null, float* out, int ldo, int B, int slot, const int32_t* len, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training-loss VALUE on device (reference models/loss.py and the compute_loss methods of models/blocks.py:313-320,
 * 369-382, 487-497, 90-106, 677-786), batched over the videos of a step.  Partial sums live in a workspace
 * ws[term][B][nchunk] (nchunk * 64 >= slot); factk_loss_combine applies the per-block formulas.  No gradients.
 * ------------------------------------------------------------------------------------------------------------------ */

/* Labels -> criterion state (MatchCriterion.set_label, loss.py:56-87): transcript[b][s] = label at the start of
 * ground-truth segment s, sweight[b][s] = cweight[transcript]; optionally inv_count[b][k] = 1 / max(#frames whose
 * (mapped) label is k, 1) and nvalid[b] = #frames with a mapped label >= 0 (InfoNCE class->frame term, loss.py:324-331).
 * cmap: optional [C] class remapping (holdout classes -> -1, blocks.py:706-725). */
int factk_label_prep(const int32_t* label, const int32_t* seg_start, const int32_t* nseg, const float* cweight,
                     const int32_t* cmap, int C, int32_t* transcript, float* sweight, int smax, float* inv_count,
                     int32_t* nvalid, int B, int slot, const int32_t* len, void* stream);

/* Matching cost (MatchCriterion.match / a2f_soft_iou, loss.py:91-136): cost[b][a][s] = -pc * softmax(aclogit)[a,
 * transcript[s]] - a2fc * IoU(a, s); attn rows are frames, or predicted segments reached through ridx[b][t].
 * overlap: scratch [B][smax][ldo].  logp != 0: aclogit rows are log-probabilities used as exp(.) (verb/noun model). */
int factk_match_cost(const float* attn, int lda, int aslot, const int32_t* ridx, const float* aclogit, int M, int C1,
                     const int32_t* transcript, const int32_t* seg_start, const int32_t* seg_len, const int32_t* nseg,
                     int smax, float pc, float a2fc, float* overlap, int ldo, float* cost, int B, int slot, int logp, void* stream);

/* part_sum[b][chunk] = sum over the chunk's frames t of  -(X[b][r][c] - lse) * w[b*w_bstride + k] / rlen[b][r]  with
 * k = tmap ? tmap[b*tmap_bstride + tgt0[b][t]] : tgt0[b][t] (frames with k < 0 are skipped), c = cols ? cols[b*cols_bstride
 * + k] : k, r = ridx ? ridx[b][t] : t, and lse = col_lse ? col_lse[b][c] : log-sum-exp of row r over the selected columns.
 * Serves frame_loss(_tdu), cross_attn_loss(_tdu) in both directions and both InfoNCE terms (loss.py:211-277, 280-341). */
int factk_loss_pick(const float* X, int ldx, int xslot, int ncol, const int32_t* cols, int cols_bstride,
                    const int32_t* ncols, const int32_t* ridx, const int32_t* rlen, const float* col_lse, int ld_lse,
                    const int32_t* tgt0, const int32_t* tmap, int tmap_bstride, const float* w, int w_bstride,
                    float* part_sum, float* part_cnt, int B, int slot, const int32_t* len, int nchunk, void* stream);

/* smooth_loss numerator (loss.py:8-19): sum over t < len-1 and c of min((logp[t+1][c] - logp[t][c])^2, 16); is_logp != 0:
 * X already holds log-probabilities (is_logit=False). */
int factk_loss_smooth(const float* X, int ldx, int ncol, float* part, int B, int slot, const int32_t* len, int nchunk,
                      int is_logp, void* stream);

/* out[b][c] = log-sum-exp over rows r < nrows[b] of X[b][r][c]; rows whose (mapped) rmask0 entry is negative are skipped. */
/* ws: factk_col_lse_ws_floats(B, xslot, ncol) floats (0: not needed; NULL is always accepted -- one CTA column per 32 columns
 * then walks all rows); with it long rows tensors are reduced in 512-row chunks across CTAs, combined in chunk order. */
size_t factk_col_lse_ws_floats(int B, int xslot, int ncol);
int factk_col_lse(const float* X, int ldx, int xslot, int ncol, const int32_t* nrows, const int32_t* rmask0,
                  const int32_t* rmap, int rmap_bstride, float* out, int ldo, int B, float* ws, void* stream);

/* action_token_loss (loss.py:196-209): out[b*out_stride] = class-weighted cross entropy of the tokens, unmatched tokens
 * labelled with the null class C1-1.  logp_mean != 0: the verb/noun model's form (blocks_SepVerbNoun.py:254-266): rows
 * are log-probabilities, mean over the tokens of weight * (-logp[label]). */
int factk_token_loss(const float* aclogit, int M, int C1, const int32_t* aind, const int32_t* sind, const int32_t* nmatch,
                     int kmax, const int32_t* transcript, int smax, const float* cweight, float* out, int out_stride,
                     int B, int logp_mean, void* stream);

/* Per-block formulas, block mean and the FACT / InfoNCE mix.  block_type: HOST array [nb] (0 input, 1 update, 2 update
 * with temporal down/up-sampling, 3 / 4 verb/noun input / update block); npred: [nb][B] predicted segment counts (types >= 2).  out[b] = {loss, fact_loss,
 * contrastive_loss, has_contrastive, block losses...}. */
int factk_loss_combine(const float* ws, int nb, const int32_t* block_type, int B, int nchunk, const int32_t* len,
                       const int32_t* npred, int C, int M, float sw, int use_clip, float fact_w, float con_w, int nseen,
                       const int32_t* nvalid, float* out, int ldo, void* stream);


/* One launch per token-side decoder layer (csrc/token_layer.cu; models/basic.py:429-452 SALayer.forward, 494-523
 * SCALayer.forward), one CTA per video, the token matrix resident in shared memory between the stages:
 *   [w_in: packed q|k|v projection of x (+ pre_qk[m][:2A], the query positions times W) -> per-head softmax(q k^T) v]
 *   (or o_in: the rows that enter out_proj, e.g. the cross attention's output) -> out_proj + b_o + x -> LayerNorm(ln1)
 *   [w_q: cq_out = LN1 rows W_q^T + b_q + pre_q]   [w_1: FFN linear1 / ReLU / linear2 + residual -> LayerNorm(ln2)]
 * and x <- the last LayerNorm's rows.  x, o_in, cq_out: fp32 [B][M][A]; biases / LayerNorm parameters / tables fp32.
 * Weights: bf16 in mma.m16n8k16 B-fragment order -- W[N][K] as [N/32][K/16][lane = 4 (n % 8) + (k % 8) / 2][(n / 8) % 4]
 * [(k / 8) % 2][k % 2] (ops.pack_token_weight).  Limits: M <= 80, A % 64 == 0 <= 256, head dim 16/32/64, ff % 64 == 0
 * (factk_token_layer_supported). */
typedef struct {
    float* x;
    int32_t B, M, A, nhead, ff;
    float eps;
    const void* w_in;
    const float* b_in;
    const float* pre_qk;
    const float* o_in;
    const void* w_o;
    const float* b_o;
    const float* ln1_w;
    const float* ln1_b;
    const void* w_q;
    const float* b_q;
    const float* pre_q;
    float* cq_out;
    const void* w_1;
    const float* b_1;
    const void* w_2;
    const float* b_2;
    const float* ln2_w;
    const float* ln2_b;
} factk_token_layer_t;
int factk_token_layer_supported(int M, int A, int nhead, int ff);
int factk_token_layer(const factk_token_layer_t* p, void* stream);
/* Debug aid: clock64 of CTA 0 at the kernel's ten stage boundaries into buf[10] (int64, device memory); NULL turns it off. */
int factk_token_layer_debug(void* buf);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step: backward-pass primitives (csrc/train.cu, train_gru.cu, train_loss.cu).  The reference leaves the backward
 * to torch autograd over its eager ops (scripts/train.py:262-264 `loss.backward()`); here every gradient is one of the
 * kernels below, or the forward GEMM itself run with transposed weights and negated tap offsets (the data gradient of a
 * Conv1d tap / Linear).  All reductions are two-stage with a fixed order: gradients are bit-reproducible.
 * ------------------------------------------------------------------------------------------------------------------ */

/* Weight gradient of one GEMM source (nn.Conv1d tap / nn.Linear; models/basic.py:138-139,158,177,182,343-347):
 *   dW[(b)][n][k] (+)= alpha * sum over valid rows t of dZ[b,t,n] * (A[b, t+row_off, k] + pos[pidx(t), k] for k < pos_d)
 * dw_bstride == 0: the videos are summed (shared weight); != 0: one [N][K] block per video (per-video "weights": the
 * attention operands).  The same contraction is "attention^T @ values" of the softmax-over-frames direction
 * (basic.py:373-379) in the training forward.  ws: factk_wgrad_ws_floats(...) floats of scratch. */
size_t factk_wgrad_ws_floats(int B, int slot, int N, int K);
int factk_wgrad(const void* dZ, int dz_dtype, int lddz, const void* A, int a_dtype, int lda, int a_slot, int row_off,
                const float* pos, int pos_ld, int pos_d, const int32_t* pos_idx, int N, int K, float* dW, int lddw,
                long long dw_bstride, float alpha, int accumulate, int B, int slot, const int32_t* len, float* ws,
                void* stream);

/* Head-batched products of nn.MultiheadAttention's forward / backward (models/basic.py:437,500,507-514; csrc/train_attn.cu),
 * one launch for all heads of all videos:
 *   C[b][h](m, n) (+)= alpha * sum_k A[b][h](m, k) * Bm[b][h](n, k),   m < M, n < N, k < K
 * Each operand is a column band of a rows tensor: (video b, head h) starts at element b * bstride + h * hstride; a
 * k-contiguous operand holds (m, k) at [m * ld + k], a k-major one at [k * ld + m]; C holds (m, n) at [m * ldc + n].
 * len_mode 1: len[b] bounds m (the frames are the rows of C); 2: len[b] bounds k (a reduction over the frames: k is split
 * across CTAs and the partials are summed in a fixed order - ws: factk_heads_mm_ws_floats(...) floats, may be NULL when 0).
 * The six products of one attention are three layouts: logits / dP (A, Bm k-contiguous, K = head dim), apply / dQ-of-rows
 * (A k-contiguous probabilities, Bm k-major), and the transposed apply (A, Bm k-major). */
typedef struct {
    const void* A;
    int32_t a_dtype, lda;
    int64_t a_bstride;
    int32_t a_hstride, a_kmajor;
    const void* Bm;
    int32_t b_dtype, ldb;
    int64_t b_bstride;
    int32_t b_hstride, b_kmajor;
    void* C;
    int32_t c_dtype, ldc;
    int64_t c_bstride;
    int32_t c_hstride, accumulate;
    int32_t M, N, K, batch, nhead, len_mode;
    const int32_t* len;
    float* ws;
    float alpha;
    int32_t reserved_;
} factk_heads_mm_t;
size_t factk_heads_mm_ws_floats(int batch, int nhead, int M, int N, int K);
int factk_heads_mm(const factk_heads_mm_t* g, void* stream);

/* The same contraction on the tensor cores (csrc/wgrad_tc.cu): bf16 dZ and A, both consumed as MN-major tcgen05 operands straight
 * from their row-major layout (TMA boxes of 64 frames x 64 channels), fp32 accumulation in tensor memory, per-chunk partial
 * tiles summed in a fixed order.  Needs N % 64 == 0, K % 64 == 0, 16-byte aligned rows, slot % 64 == 0, and ZERO rows in
 * dZ at [len[b], slot).  ws: factk_wgrad_tc_ws_floats(...) floats. */
int factk_wgrad_tc_supported(int dz_dtype, int lddz, int a_dtype, int lda, int N, int K, int slot);
size_t factk_wgrad_tc_ws_floats(int B, int slot, int N, int K);
int factk_wgrad_tc(const void* dZ, int lddz, const void* A, int lda, int a_slot, int row_off, int N, int K, float* dW,
                   int lddw, long long dw_bstride, float alpha, int accumulate, int B, int slot, const int32_t* len,
                   float* ws, void* stream);

/* out[(b)][n] (+)= alpha * sum over valid rows of X[b,t,n] (* Y[b,t,n] when Y != NULL): bias gradients, softmax column
 * terms.  out_bstride == 0 sums the videos too. */
size_t factk_colsum_ws_floats(int B, int slot, int N);
int factk_colsum(const void* X, int x_dtype, int ldx, const void* Y, int y_dtype, int ldy, int N, float* out,
                 long long out_bstride, float alpha, int accumulate, int B, int slot, const int32_t* len, float* ws,
                 void* stream);

/* Elementwise passes over the valid rows.  op: 0 Y = X * (R > 0) (ReLU backward, R = the ReLU output); 1 Y += alpha * X;
 * 2 Y = dropout(X) (nn.Dropout, p, mask = hash(seed, site, element) -- identical in forward and backward, nothing stored);
 * 3 Y = channel dropout (nn.Dropout2d over (1,D,T): one decision per (video, channel), blocks.py:28,614-617);
 * 4 Y = alpha * X; 5 Y = X * R; 6 Y = X + R; 7 Y = max(X, 0); 8 Y = X * R[row] (one value per row: the time mask of
 * basic.time_mask); 9 Y = X + R[ridx ? ridx[b,t] : t] with R one table shared by the videos (add_positional_encoding with the
 * sinusoid table, rows by frame index or by segment centre, basic.py:313-320, blocks.py:454-455).  seed_ptr (optional, device memory) is added to seed, so a captured CUDA graph draws new masks on every replay.  x_slot: row slots per video of X (< 0: slot; 0 broadcasts one
 * table over the videos -- the learned action_query / positional terms of add_positional_encoding, basic.py:313-320). */
#define FACTK_EW_RELU_BWD 0
#define FACTK_EW_AXPY 1
#define FACTK_EW_DROPOUT 2
#define FACTK_EW_DROPOUT_CH 3
#define FACTK_EW_COPY 4
#define FACTK_EW_MUL 5
#define FACTK_EW_ADD 6
#define FACTK_EW_RELU 7
#define FACTK_EW_ROWSCALE 8
#define FACTK_EW_ADDTAB 9
int factk_rows_elementwise(int op, const void* X, int x_dtype, int ldx, const void* R, int r_dtype, int ldr, void* Y,
                           int y_dtype, int ldy, int N, int B, int slot, const int32_t* len, float alpha, float p,
                           unsigned long long seed, unsigned site, int x_slot, const unsigned long long* seed_ptr,
                           const int32_t* ridx, void* stream);

/* dst[b][c][r] = src[b][r][c], fp32 (token-sized operands of the attention gradients). */
int factk_transpose(const float* src, int lds, long long src_bstride, float* dst, int ldd, long long dst_bstride, int R,
                    int Ccols, int B, void* stream);

/* Block.process_feature backward (blocks.py:195-202): Y = the spliced rows [feat | softmax(logits)], dY their gradient
 * (may be NULL), dCl the gradient of the raw logits returned beside them (may be NULL) -> dX, the gradient of the rows
 * before the splice. */
int factk_splice_bwd(const void* Y, int y_dtype, int ldy, const void* dY, int dy_dtype, int lddy, const float* dCl,
                     int lddc, void* dX, int dx_dtype, int lddx, int H, int C, int B, int slot, const int32_t* len,
                     void* stream);

/* Row softmax backward (X2Y_map a2f direction, token self attention): dL (+)= P * (dP - sum_m P dP). */
int factk_row_softmax_bwd(const float* P, int ldp, const float* dP, int lddp, float* dL, int lddl, int M, int accumulate,
                          int B, int slot, const int32_t* len, void* stream);

/* LayerNorm backward: y = relu?(LN(x (+ r)) * w + b).  dV (+)= gradient of v = x + r; dw, db += affine gradients. */
size_t factk_layernorm_bwd_ws_floats(int B, int slot, int E);
int factk_layernorm_bwd(const void* X, int x_dtype, int ldx, const void* R, int r_dtype, int ldr, const float* w,
                        const float* b, float eps, int relu, const void* dY, int dy_dtype, int lddy, void* dV, int dv_dtype,
                        int lddv, int accumulate, float* dw, float* db, int B, int slot, const int32_t* len, int E,
                        float* ws, void* stream);

/* F.normalize backward (blocks.py:174). */
int factk_l2norm_bwd(const void* X, int x_dtype, int ldx, const void* dY, int dy_dtype, int lddy, void* dX, int dx_dtype,
                     int lddx, int B, int slot, const int32_t* len, int E, float eps, void* stream);

/* Softmax over ROWS with the normalised attention kept (training forward of X2Y_map f2a, basic.py:373-376, and of the
 * SCALayer cross attention per head, basic.py:507-514) and its backward dL (+)= scale * P * (dP - colsum(P dP)). */
size_t factk_col_softmax_train_ws_floats(int B, int slot, int M);
int factk_col_softmax(const void* L, int l_dtype, int ldl, void* P, int p_dtype, int ldp, int M, float scale, int B, int slot,
                      const int32_t* len, float* ws, void* stream);
/* ws: factk_colsum_ws_floats(B, slot, M) + B * M floats */
int factk_col_softmax_bwd(const void* P, int p_dtype, int ldp, const void* dP, int dp_dtype, int lddp, void* dL, int dl_dtype, int lddl,
                          int M, float scale, int accumulate, int B, int slot, const int32_t* len, float* ws, void* stream);

/* Segment mean backward / gathered pre-activation backward (basic.py:615-625, blocks.py:439-447):
 * reduce: out[b][s] (+)= (mean ? 1/len : 1) * sum of the segment's frame rows; expand: out[b][t] (+)= (inv_len ? 1/len : 1) *
 * seg[b][seg_label[t]]. */
int factk_segment_reduce(const void* X, int x_dtype, int ldx, void* out, int o_dtype, int ldo, const int32_t* seg_start,
                         const int32_t* seg_len, const int32_t* nseg, int B, int slot, int E, int mean, int accumulate,
                         void* stream);
int factk_segment_expand(const void* seg, int s_dtype, int lds, const int32_t* seg_label, const int32_t* seg_len, void* out,
                         int o_dtype, int ldo, int B, int slot, const int32_t* len, int E, int inv_len, int accumulate,
                         void* stream);

/* BPTT of factk_gru_bidir (nn.GRU backward, blocks.py:401,432), one layer, one 4-CTA cluster per (video, direction).
 * gi, gh fp32 [B][slot][6*Hh] (gh = h_{t-1} W_hh^T + b_hh for every step: one GEMM over the saved hidden states with row
 * offset -1 / +1); hout = the layer output before any ReLU; dout its gradient -> dgi, dgh fp32 [B][slot][6*Hh]. */
int factk_gru_bwd(const float* gi, const float* gh, const void* hout, int h_dtype, int ldh, const void* dout, int do_dtype,
                  int lddo, const float* w_hh_f, const float* w_hh_b, int Hh, float* dgi, float* dgh, int B, int slot,
                  const int32_t* nseg, void* stream);
/* Debug aid: clock64 at seven points of the first 64 steps of the Hh = 256 BPTT kernel into buf[64][8] (int64, device); NULL: off. */
int factk_gru_bwd_debug(void* buf);

/* Gradients of the training loss w.r.t. the logit tensors it reads (csrc/train_loss.cu; reference: autograd through
 * models/loss.py:8-19,196-341 and the compute_loss methods of models/blocks.py:313-320,369-382,487-497,677-786).  Each call
 * ADDS coef[b] * d(term)/dX into dX (same layout as X); coef carries block mean, FACT / InfoNCE mix and the batch mean. */
/* frame_loss (seg_start == NULL, rows = frames, nrows = len) / frame_loss_tdu (rows = predicted segments, nrows = nseg). */
int factk_loss_grad_ce_rows(const float* X, int ldx, int C, float* dX, int lddx, const int32_t* label, const float* cweight,
                            const int32_t* seg_start, const int32_t* seg_len, const int32_t* nrows, const float* coef,
                            int B, int slot, void* stream);
/* smooth_loss on logits; mult = Loss.sw. */
int factk_loss_grad_smooth(const float* X, int ldx, int C, float* dX, int lddx, const int32_t* len, const float* coef,
                           float mult, int B, int slot, void* stream);
/* action_token_loss. */
int factk_loss_grad_token(const float* aclogit, int M, int C1, float* dX, const int32_t* aind, const int32_t* sind,
                          const int32_t* nmatch, int kmax, const int32_t* transcript, int smax, const float* cweight,
                          const float* coef, int B, void* stream);
/* cross_attn_loss(_tdu): mode 0 = a2f direction (softmax over the matched tokens of a row; mult[b][m] = multiplicity of
 * token m among the matches), mode 1 = f2a direction (softmax over the rows of a column; col_lse from factk_col_lse, colmass
 * scratch [B][M]).  colmap / wmap [B][smax]: matched token and weight per ground-truth segment.  seg_* NULL: rows = frames. */
int factk_loss_grad_xattn(int mode, const float* X, int ldx, int M, float* dX, int lddx, const int32_t* gseg,
                          const int32_t* gstart, const int32_t* glen, const int32_t* gn, const int32_t* colmap,
                          const float* wmap, int smax, const float* mult, float* colmass, const float* col_lse, int ld_lse,
                          const int32_t* seg_label, const int32_t* seg_start, const int32_t* seg_len, const int32_t* nrows,
                          const float* coef, int B, int slot, void* stream);
/* infonce_contrastive_loss on sim = emb . text^T / temp over the seen classes (cmap[c] >= 0). */
int factk_loss_grad_infonce(const float* sim, int lds, int C, float* dS, int ldds, const int32_t* label, const int32_t* cmap,
                            float* count, const float* col_lse, int ld_lse, const int32_t* nvalid, int nseen,
                            const int32_t* len, const float* coef, int B, int slot, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FACTK_H_ */
