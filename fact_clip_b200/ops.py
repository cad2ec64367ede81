"""Tensor-level wrappers over the factk C ABI (include/factk.h).

Every function takes CUDA torch tensors (torch is only the allocator / stream provider), launches the
hand-written kernels on the current stream and returns nothing (outputs are written in place).
Layout: "rows" tensors are [B, slot, ld] with ``len`` (int32 [B], device) valid rows per video.
"""
import ctypes
import os

import torch

from . import _lib as L


COUNTERS = {'launches': 0}
TIMER = None
PROF_SHAPES = os.environ.get('FACTK_PROF_SHAPES') == '1'        # profiling aid: key SIMT GEMM / wgrad timings by shape


class KernelTimer:
    """CUDA-event timing of tagged launches on the launching stream (used by bench.py for the roofline line)."""

    def __init__(self, tags):
        self.tags = None if tags is None else set(tags)
        self.pairs = {} if tags is None else {t: [] for t in tags}

    def collect(self, skip_steps, steps):
        torch.cuda.synchronize()
        out = {}
        for t, pairs in self.pairs.items():
            per_step = len(pairs) // max(skip_steps + steps, 1)
            use = pairs[skip_steps * per_step:]
            out[t] = dict(ms=sum(a.elapsed_time(b) for a, b in use), n=len(use))
        return out


def _call(name, tag, *args):
    """Launch through the C ABI; when a KernelTimer is active, bracket the launch with CUDA events."""
    t = TIMER
    key = tag or name
    if t is not None and (t.tags is None or key in t.tags):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.call(name, *args)
        e1.record()
        t.pairs.setdefault(key, []).append((e0, e1))
    else:
        L.call(name, *args)


def S(A, W, K=None, off=0, gather=None, pos=None, pos_d=None, pos_idx=None):
    """One GEMM operand: rows of ``A`` ([B|1, a_slot, lda]) times ``W^T`` (W: [N,K] shared or [B,N,K] per video)."""
    return dict(A=A, W=W, K=K, off=off, gather=gather, pos=pos, pos_d=pos_d, pos_idx=pos_idx)


def _row_ld(t):
    assert t.stride(-1) == 1, 'rows tensors must be contiguous in the last dimension'
    return t.stride(-2)


def gemm(srcs, N, out, len=None, bias=None, alpha=1.0, relu=False, res=None, tag=None, tc=False, pre=None, pre_idx=None):
    g = L.Gemm()
    B, slot = out.shape[0], out.shape[1]
    g.B, g.slot, g.N, g.nsrc = B, slot, N, len_(srcs)
    g.len = L.ptr(len)
    for i, s in enumerate(srcs):
        A, W = s['A'], s['W']
        x = g.src[i]
        x.A, x.a_dtype, x.lda = A.data_ptr(), L.dt(A), _row_ld(A)
        x.a_slot = 0 if (A.shape[0] == 1 and B > 1) else A.stride(0) // _row_ld(A)
        x.K = s['K'] if s['K'] is not None else W.shape[-1]
        x.row_off = s['off']
        x.gather = L.ptr(s['gather'])
        if s['pos'] is not None:
            x.pos, x.pos_ld = s['pos'].data_ptr(), s['pos'].stride(0)
            x.pos_d = s['pos_d'] if s['pos_d'] is not None else s['pos'].shape[-1]
            x.pos_idx = L.ptr(s['pos_idx'])
        assert W.stride(-1) == 1
        x.W, x.ldw, x.w_dtype = W.data_ptr(), W.stride(-2), L.dt(W)
        x.w_bstride = W.stride(0) if W.dim() == 3 else 0
    if bias is not None:
        g.bias = bias.data_ptr()
        g.bias_bstride = bias.stride(0) if bias.dim() == 2 else 0
    g.alpha, g.relu = float(alpha), int(relu)
    if pre is not None:
        g.pre, g.pre_dtype, g.ldpre = pre.data_ptr(), L.dt(pre), _row_ld(pre)
        g.pre_bstride = pre.stride(0) if pre.dim() == 3 else 0
        g.pre_idx = L.ptr(pre_idx)
    if res is not None:
        g.res, g.res_dtype, g.ldres = res.data_ptr(), L.dt(res), _row_ld(res)
    g.Y, g.y_dtype, g.ldy = out.data_ptr(), L.dt(out), _row_ld(out)
    if PROF_SHAPES and not tc:
        tag = f"{tag or 'gemm_simt'}[{B}x{slot} N={N} K={'+'.join(str(g.src[i].K) for i in range(g.nsrc))} {str(srcs[0]['A'].dtype)[6:]}>{str(out.dtype)[6:]}]"
    _call('factk_gemm_tc' if tc else 'factk_gemm', tag or ('gemm_tc' if tc else 'gemm_simt'), g, L.stream())
    COUNTERS['launches'] += 1


def gemm_pair_ok(A, W, N, out, pre=None):
    """True when factk_gemm_pair can run y = act(A W^T + bias + pre): bf16 rows / weights / output, aligned, K % 64, N % 128."""
    bf = torch.bfloat16
    if not (A.dtype == W.dtype == out.dtype == bf and W.dim() == 2 and A.dim() == 3 and A.shape[0] == out.shape[0]):
        return False
    K = W.shape[1]
    if K % 64 or N % 128 or W.shape[0] != N or A.stride(-1) != 1 or W.stride(-1) != 1 or out.stride(-1) != 1:
        return False
    lda, ldw, ldy = A.stride(-2), W.stride(-2), out.stride(-2)
    if lda % 8 or ldw % 8 or ldy % 8 or A.data_ptr() % 16 or W.data_ptr() % 16 or out.data_ptr() % 16:
        return False
    if A.stride(0) % lda or A.stride(0) // lda < out.shape[1] or out.stride(0) != out.shape[1] * ldy:
        return False
    if pre is not None and (pre.dtype != torch.float32 or pre.stride(-1) != 1 or pre.stride(-2) % 4 or pre.data_ptr() % 16):
        return False
    return True


def gemm_pair(A, W, N, out, len=None, bias=None, relu=False, pre=None, pre_idx=None, tag=None):
    B, slot = out.shape[0], out.shape[1]
    lda = A.stride(-2)
    pre_bstride = pre.stride(0) if (pre is not None and pre.dim() == 3) else 0
    COUNTERS['launches'] += 1
    _call('factk_gemm_pair', tag or 'gemm_pair', A.data_ptr(), lda, A.stride(0) // lda, W.data_ptr(), W.stride(-2), W.shape[1], N,
          L.ptr(bias), L.ptr(pre), pre.stride(-2) if pre is not None else 0, pre_bstride, L.ptr(pre_idx), int(relu),
          out.data_ptr(), out.stride(-2), B, slot, L.ptr(len), L.stream())


def tcn_layer_supported(F):
    return bool(L.load().factk_tcn_layer_supported(int(F)))


def tcn_layer(x, y, w3, b3, w1, b1, dilation, len=None, cta_group=2):
    """Fused dilated residual layer (conv3 + ReLU + 1x1 + residual) on bf16 rows [B, slot, F]."""
    B, slot, F = x.shape
    assert x.dtype == y.dtype == w3.dtype == w1.dtype == torch.bfloat16 and x.is_contiguous() and y.is_contiguous()
    assert w3.is_contiguous() and w1.is_contiguous() and tuple(w3.shape) == (3, F, F) and tuple(w1.shape) == (F, F)
    COUNTERS['launches'] += 1
    _call('factk_tcn_layer', 'tcn_layer', x.data_ptr(), y.data_ptr(), w3.data_ptr(), b3.data_ptr(), w1.data_ptr(),
          b1.data_ptr(), B, slot, F, int(dilation), L.ptr(len), int(cta_group), L.stream())


len_ = len   # the builtin (``len`` is also a keyword argument name in this module)


def softmax_splice(x, C, clogit, pred=None, len=None, H=None):
    B, slot = x.shape[0], x.shape[1]
    H = x.shape[-1] if H is None else H
    COUNTERS['launches'] += 1
    _call('factk_softmax_splice', None, x.data_ptr(), L.dt(x), B, slot, L.ptr(len), _row_ld(x), H, C,
           clogit.data_ptr(), L.ptr(pred), L.stream())


def layernorm(x, w, b, out, res=None, eps=1e-5, relu=False, len=None, E=None):
    B, slot = x.shape[0], x.shape[1]
    E = x.shape[-1] if E is None else E
    COUNTERS['launches'] += 1
    _call('factk_layernorm', None, x.data_ptr(), L.dt(x), _row_ld(x),
           L.ptr(res), L.dt(res) if res is not None else 0, _row_ld(res) if res is not None else 0,
           w.data_ptr(), b.data_ptr(), float(eps), int(relu), out.data_ptr(), L.dt(out), _row_ld(out),
           B, slot, L.ptr(len), E, L.stream())


def l2norm(x, out, eps=1e-12, len=None):
    B, slot, E = x.shape
    COUNTERS['launches'] += 1
    _call('factk_l2norm', None, x.data_ptr(), L.dt(x), _row_ld(x), out.data_ptr(), L.dt(out), _row_ld(out),
           B, slot, L.ptr(len), E, float(eps), L.stream())


def row_softmax(logit, out, M, scale=1.0, len=None, out16=None):
    B, slot = logit.shape[0], logit.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_row_softmax', None, logit.data_ptr(), _row_ld(logit), out.data_ptr(), _row_ld(out), B, slot,
           L.ptr(len), M, float(scale), L.ptr(out16), _row_ld(out16) if out16 is not None else 0,
           out16.shape[-1] if out16 is not None else 0, L.stream())


def mha_tokens(q, k, v, out, nhead, tf32=False):
    B, M = q.shape[0], q.shape[1]
    E = out.shape[-1]
    assert _row_ld(q) == _row_ld(k) == _row_ld(v)
    COUNTERS['launches'] += 1
    _call('factk_mha_tokens', None, q.data_ptr(), k.data_ptr(), v.data_ptr(), _row_ld(q), out.data_ptr(), _row_ld(out),
           B, M, nhead, E // nhead, int(tf32), L.stream())


def attn_rows_ws(B, slot, M, nhead, dh):
    return int(L.load().factk_attn_rows_ws_floats(B, slot, M, nhead, dh))


def attn_rows(q, kx, vx, out, nhead, ws, len=None):
    B, M, E = q.shape[0], q.shape[1], out.shape[-1]
    slot = kx.shape[1]
    assert _row_ld(kx) == _row_ld(vx) and kx.dtype == vx.dtype
    COUNTERS['launches'] += 2
    _call('factk_attn_rows', None, q.data_ptr(), _row_ld(q), kx.data_ptr(), vx.data_ptr(), L.dt(kx), _row_ld(kx),
           out.data_ptr(), _row_ld(out), B, slot, L.ptr(len), M, nhead, E // nhead, ws.data_ptr(), L.stream())


def col_softmax_ws(B, slot, M, E):
    return int(L.load().factk_col_softmax_ws_floats(B, slot, M, E))


def col_softmax_apply(logit, x, out, M, ws, attn=None, len=None, E=None):
    B, slot = logit.shape[0], logit.shape[1]
    E = x.shape[-1] if E is None else E
    COUNTERS['launches'] += 4
    _call('factk_col_softmax_apply', None, logit.data_ptr(), _row_ld(logit), x.data_ptr(), L.dt(x), _row_ld(x),
           out.data_ptr(), _row_ld(out), L.ptr(attn), _row_ld(attn) if attn is not None else 0,
           B, slot, L.ptr(len), M, E, ws.data_ptr(), L.stream())


def a2f_fused_ok(M, H, F, slot):
    return bool(L.load().factk_a2f_fused_supported(int(M), int(H), int(F), int(slot)))


def a2f_fused(rows, kt, cb, wy, vt, bias, out, M, logit=None, attn=None, len=None):
    """Fused X2Y_map (a2f direction) on bf16 rows [B, slot, H]; kt [B, M, H] bf16, cb [B, M] fp32, wy [F, H] bf16, vt [B, F, Kp] bf16."""
    B, slot, H = rows.shape
    F = wy.shape[0]
    bf = torch.bfloat16
    assert rows.dtype == kt.dtype == wy.dtype == vt.dtype == out.dtype == bf and cb.dtype == torch.float32
    assert cb.dim() == 2 and cb.stride(1) == 1
    ref = logit if logit is not None else attn
    COUNTERS['launches'] += 1
    _call('factk_a2f_fused', 'a2f_fused', rows.data_ptr(), _row_ld(rows), kt.data_ptr(), _row_ld(kt), kt.stride(0), cb.data_ptr(), cb.stride(0),
          wy.data_ptr(), _row_ld(wy), vt.data_ptr(), _row_ld(vt), vt.stride(0), bias.data_ptr(), out.data_ptr(), _row_ld(out),
          L.ptr(logit), L.ptr(attn), _row_ld(ref) if ref is not None else 0, B, slot, L.ptr(len), M, H, F, L.stream())


def f2a_fused_ok(M, H, slot):
    return bool(L.load().factk_f2a_fused_supported(int(M), int(H), int(slot)))


def f2a_fused_ws(B, slot, M, H):
    return int(L.load().factk_f2a_fused_ws_floats(B, slot, M, H))


def f2a_fused(rows, qt, out, M, ws, len=None):
    """Fused X2Y_map (f2a direction): rows [B, slot, H] bf16, qt [B, M, H] bf16 -> out [B, M, H] fp32 (softmax over the rows of each
    video, weighted row sum); one tcgen05 launch + the split combine."""
    B, slot, H = rows.shape
    assert rows.dtype == qt.dtype == torch.bfloat16 and out.dtype == ws.dtype == torch.float32
    assert rows.stride(0) == slot * _row_ld(rows)
    COUNTERS['launches'] += 2
    _call('factk_f2a_fused', 'f2a_fused', rows.data_ptr(), _row_ld(rows), qt.data_ptr(), _row_ld(qt), qt.stride(0), out.data_ptr(), _row_ld(out),
          B, slot, L.ptr(len), M, H, ws.data_ptr(), L.stream())


def token_layer_ok(M, A, nhead, ff):
    return bool(L.load().factk_token_layer_supported(M, A, nhead, ff))


def pack_token_weight(W):
    """[N, K] weight -> bf16 in mma.m16n8k16 B-fragment order (include/factk.h factk_token_layer_t): per (32-column task, k-step of
    16, lane) the eight registers the lane feeds to the four n8 tiles of the task."""
    N, K = W.shape
    assert N % 32 == 0 and K % 16 == 0
    w = W.to(torch.bfloat16).reshape(N // 32, 4, 8, K // 16, 2, 4, 2)        # [task, tile, n % 8, k-step, k half, (k % 8) / 2, k % 2]
    return w.permute(0, 3, 2, 5, 1, 4, 6).contiguous().view(-1)


def token_layer(x, nhead, w_o, b_o, ln1_w, ln1_b, w_in=None, b_in=None, pre_qk=None, o_in=None, w_q=None, b_q=None, pre_q=None, cq_out=None,
                ffn=None, eps=1e-5):
    """One launch per token-side decoder (half-)layer (csrc/token_layer.cu).  x: [B, M, A] fp32 contiguous, updated in place;
    weights packed by pack_token_weight; ffn = (w_1, b_1, w_2, b_2, ln2_w, ln2_b, ff) or None."""
    B, M, A = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32
    g = L.TokenLayer()
    g.x, g.B, g.M, g.A, g.nhead, g.eps = x.data_ptr(), B, M, A, nhead, float(eps)
    g.w_in, g.b_in, g.pre_qk = L.ptr(w_in), L.ptr(b_in), L.ptr(pre_qk)
    if o_in is not None:
        assert o_in.is_contiguous() and o_in.dtype == torch.float32 and o_in.shape == x.shape
    g.o_in = L.ptr(o_in)
    g.w_o, g.b_o, g.ln1_w, g.ln1_b = w_o.data_ptr(), b_o.data_ptr(), ln1_w.data_ptr(), ln1_b.data_ptr()
    if cq_out is not None:
        assert cq_out.is_contiguous() and cq_out.dtype == torch.float32 and cq_out.shape == x.shape
    g.w_q, g.b_q, g.pre_q, g.cq_out = L.ptr(w_q), L.ptr(b_q), L.ptr(pre_q), L.ptr(cq_out)
    if ffn is not None:
        w1, b1, w2, b2, l2w, l2b, ff = ffn
        g.w_1, g.b_1, g.w_2, g.b_2, g.ln2_w, g.ln2_b, g.ff = w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), l2w.data_ptr(), l2b.data_ptr(), ff
    COUNTERS['launches'] += 1
    _call('factk_token_layer', 'token_layer', g, L.stream())


def tdu_segment(pred, seg_label, seg_start, seg_len, seg_center, nseg, len=None):
    B, slot = pred.shape
    COUNTERS['launches'] += 1
    _call('factk_tdu_segment', None, pred.data_ptr(), B, slot, L.ptr(len), seg_label.data_ptr(), seg_start.data_ptr(),
           seg_len.data_ptr(), seg_center.data_ptr(), nseg.data_ptr(), L.stream())


def segment_mean_ws(B, slot, E):
    return int(L.load().factk_segment_mean_ws_floats(B, slot, E))


def segment_mean(x, seg, seg_label, seg_start, seg_len, nseg, E=None, ws=None):
    B, slot = x.shape[0], x.shape[1]
    E = x.shape[-1] if E is None else E
    COUNTERS['launches'] += 2 if ws is not None else 1
    _call('factk_segment_mean', None, x.data_ptr(), L.dt(x), _row_ld(x), seg.data_ptr(), L.dt(seg), _row_ld(seg),
           seg_label.data_ptr(), seg_start.data_ptr(), seg_len.data_ptr(), nseg.data_ptr(), B, slot, E, L.ptr(ws), L.stream())


def gru_bidir(gi, w_hh_f, b_hh_f, w_hh_b, b_hh_b, out, nseg, relu=True, mma=False, order_ws=None):
    """mma=True: tensor-core kernel batched over videos (bf16 weights / exchanged state; Hh = 256 only); with order_ws
    (int32 [B] scratch) the videos are grouped by decreasing segment count on the device first."""
    B, slot = gi.shape[0], gi.shape[1]
    Hh = w_hh_f.shape[1]
    COUNTERS['launches'] += 1
    args = (gi.data_ptr(), w_hh_f.data_ptr(), b_hh_f.data_ptr(), w_hh_b.data_ptr(), b_hh_b.data_ptr(),
            Hh, out.data_ptr(), L.dt(out), _row_ld(out), int(relu), B, slot, nseg.data_ptr())
    if mma and order_ws is not None:
        COUNTERS['launches'] += 1
        _call('factk_gru_bidir_mma_sorted', 'factk_gru_bidir', *args, order_ws.data_ptr(), L.stream())
    else:
        _call('factk_gru_bidir_mma' if mma else 'factk_gru_bidir', 'factk_gru_bidir', *args, L.stream())


def gather_rows(src, idx, out, E, len=None):
    B, slot = out.shape[0], out.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_gather_rows', None, src.data_ptr(), _row_ld(src), src.shape[1], idx.data_ptr(), out.data_ptr(),
           _row_ld(out), B, slot, L.ptr(len), E, L.stream())


def fuse_eval(action_clogit, attn, flogit, weight, pred, M, C, seg_label=None, len=None, f_logp=False):
    B, slot = flogit.shape[0], flogit.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_fuse_eval', None, L.ptr(action_clogit), L.ptr(attn), _row_ld(attn) if attn is not None else 0,
           attn.shape[1] if attn is not None else 0, L.ptr(seg_label), flogit.data_ptr(), _row_ld(flogit),
           float(weight), pred.data_ptr(), B, slot, L.ptr(len), M, C, int(f_logp), L.stream())


def vn_splice(x, n1, n2, clogit, vids=None, nids=None, pred=None, len=None, H=None):
    """Verb/noun process_feature in place on x [B, slot, ld]; with pred: segmentation argmax over the action table."""
    B, slot = x.shape[0], x.shape[1]
    H = x.shape[-1] if H is None else H
    COUNTERS['launches'] += 1
    _call('factk_vn_splice', None, x.data_ptr(), L.dt(x), B, slot, L.ptr(len), _row_ld(x), H, n1, n2, clogit.data_ptr(),
          L.ptr(vids), L.ptr(nids), vids.numel() if vids is not None else 0, L.ptr(pred), L.stream())


def vn_combine(clogit, k1, k2, vids, nids, out, with_null=False, len=None):
    B, slot = clogit.shape[0], clogit.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_vn_combine', None, clogit.data_ptr(), _row_ld(clogit), k1, k2, vids.data_ptr(), nids.data_ptr(), vids.numel(),
          int(with_null), out.data_ptr(), _row_ld(out), B, slot, L.ptr(len), L.stream())


def fuse_eval_transcript(attn, flogit, weight, transcript, ntr, pred, C, seg_label=None, len=None):
    """transcript: int32 [B, ldt]; ntr: int32 [B] entries per video (FACT.trans models)."""
    B, slot = flogit.shape[0], flogit.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_fuse_eval_transcript', None, attn.data_ptr(), _row_ld(attn), attn.shape[1], L.ptr(seg_label), flogit.data_ptr(),
          _row_ld(flogit), float(weight), transcript.data_ptr(), transcript.stride(0), ntr.data_ptr(), pred.data_ptr(), B, slot,
          L.ptr(len), C, L.stream())


def embed_tokens(embed, transcript, pe, out):
    """out[n] = embed[transcript[n]] + pe[n]  (fp32; transcript int32 [N])."""
    N, A = out.shape[-2], out.shape[-1]
    COUNTERS['launches'] += 1
    _call('factk_embed_tokens', None, embed.data_ptr(), embed.stride(0), transcript.data_ptr(), pe.data_ptr(), pe.stride(0),
          out.data_ptr(), _row_ld(out), N, A, L.stream())


# ---------------------------------------------------------------------------------------------- loss value (loss.cu)
LOSS_CHUNK = 64


def loss_nchunk(slot):
    return (slot + LOSS_CHUNK - 1) // LOSS_CHUNK


def label_prep(label, seg_start, nseg, cweight, C, transcript, sweight, len, cmap=None, inv_count=None, nvalid=None):
    B, slot = label.shape
    COUNTERS['launches'] += 1
    _call('factk_label_prep', None, label.data_ptr(), seg_start.data_ptr(), nseg.data_ptr(), cweight.data_ptr(), L.ptr(cmap), C,
          transcript.data_ptr(), sweight.data_ptr(), transcript.shape[1], L.ptr(inv_count), L.ptr(nvalid), B, slot,
          len.data_ptr(), L.stream())


def match_cost(attn, aclogit, transcript, seg_start, seg_len, nseg, pc, a2fc, overlap, cost, M, ridx=None, logp=False):
    """attn [B, aslot, >=M] (frame rows, or predicted-segment rows with ridx), aclogit [B, M, C+1] -> cost [B, M, smax]."""
    B, slot = seg_start.shape
    COUNTERS['launches'] += 2
    _call('factk_match_cost', None, attn.data_ptr(), _row_ld(attn), attn.shape[1], L.ptr(ridx), aclogit.data_ptr(), M,
          aclogit.shape[2], transcript.data_ptr(), seg_start.data_ptr(), seg_len.data_ptr(), nseg.data_ptr(),
          transcript.shape[1], float(pc), float(a2fc), overlap.data_ptr(), _row_ld(overlap), cost.data_ptr(), B, slot, int(logp),
          L.stream())


def loss_pick(X, ncol, tgt0, part, len, cols=None, ncols=None, ridx=None, rlen=None, col_lse=None, tmap=None, w=None,
              part_cnt=None):
    """See factk_loss_pick (include/factk.h).  cols / tmap / w: [B, n] per video, or 1-D shared tables."""
    B, slot = tgt0.shape
    bs = lambda t: 0 if t is None or t.dim() == 1 else t.stride(0)
    COUNTERS['launches'] += 1
    _call('factk_loss_pick', None, X.data_ptr(), _row_ld(X), X.shape[1], ncol, L.ptr(cols), bs(cols), L.ptr(ncols), L.ptr(ridx),
          L.ptr(rlen), L.ptr(col_lse), col_lse.stride(0) if col_lse is not None else 0, tgt0.data_ptr(), L.ptr(tmap), bs(tmap),
          L.ptr(w), bs(w), part.data_ptr(), L.ptr(part_cnt), B, slot, len.data_ptr(), part.shape[-1], L.stream())


def loss_smooth(X, ncol, part, len, is_logp=False):
    B, slot = X.shape[0], X.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_loss_smooth', None, X.data_ptr(), _row_ld(X), ncol, part.data_ptr(), B, slot, len.data_ptr(), part.shape[-1],
          int(is_logp), L.stream())


def col_lse(X, ncol, nrows, out, rmask0=None, rmap=None):
    B = X.shape[0]
    n = L.load().factk_col_lse_ws_floats(B, X.shape[1], ncol)
    ws = _ws(X.device, n) if n else None
    COUNTERS['launches'] += 2 if n else 1
    _call('factk_col_lse', None, X.data_ptr(), _row_ld(X), X.shape[1], ncol, nrows.data_ptr(), L.ptr(rmask0), L.ptr(rmap),
          0 if rmap is None or rmap.dim() == 1 else rmap.stride(0), out.data_ptr(), out.stride(0), B, L.ptr(ws), L.stream())


def token_loss(aclogit, aind, sind, nmatch, transcript, cweight, out, logp_mean=False):
    """out: [B, nchunk] slice of the workspace; the value lands in out[:, 0]."""
    B, M, C1 = aclogit.shape
    COUNTERS['launches'] += 1
    _call('factk_token_loss', None, aclogit.data_ptr(), M, C1, aind.data_ptr(), sind.data_ptr(), nmatch.data_ptr(), aind.shape[1],
          transcript.data_ptr(), transcript.shape[1], cweight.data_ptr(), out.data_ptr(), out.stride(0), B, int(logp_mean), L.stream())


def loss_combine(ws, block_types, len, npred, C, M, sw, out, use_clip=False, fact_w=1.0, con_w=0.0, nseen=0, nvalid=None):
    nb, B, nchunk = len_(block_types), ws.shape[1], ws.shape[2]
    bt = (ctypes.c_int32 * nb)(*block_types)
    COUNTERS['launches'] += 1
    _call('factk_loss_combine', None, ws.data_ptr(), nb, ctypes.cast(bt, ctypes.c_void_p), B, nchunk, len.data_ptr(), L.ptr(npred),
          C, M, float(sw), int(use_clip), float(fact_w), float(con_w), int(nseen), L.ptr(nvalid), out.data_ptr(), out.stride(0),
          L.stream())


def transpose_rows(src, src_bstride, dst, len, D):
    """dst[b, t, :D] = src_b[:, t] for channel-major host features copied to the device as they are (staging.py)."""
    B, slot = dst.shape[0], dst.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_transpose_rows', None, src.data_ptr(), int(src_bstride), dst.data_ptr(), L.dt(dst), _row_ld(dst), B, slot, D,
          len.data_ptr(), L.stream())


# ---------------------------------------------------------------------------------------------- training step (train*.cu)
EW_RELU_BWD, EW_AXPY, EW_DROPOUT, EW_DROPOUT_CH, EW_COPY, EW_MUL, EW_ADD, EW_RELU, EW_ROWSCALE, EW_ADDTAB = range(10)
_WS = {}


def _ws(dev, n):
    """Scratch of at least n floats on ``dev`` (grown geometrically, reused by every reduction of the backward pass --
    launches on one stream are ordered, so consecutive users cannot overlap)."""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)         # one scratch per stream: users on different streams may overlap
    t = _WS.get(key)
    if t is None or t.numel() < n:
        t = _WS[key] = torch.empty(max(int(n * 1.25), 1 << 20), dtype=torch.float32, device=dev)
    return t


def wgrad(dz, a, N, K, dw, off=0, len=None, alpha=1.0, accumulate=True, pos=None, pos_d=None, pos_idx=None, per_video=False, tc=False):
    """dw[(b)][n][k] (+)= alpha * sum_rows dz[b,t,n] * (a[b,t+off,k] + pos).  dz / a: rows tensors; dw: [N, K] view (unit
    stride in k), or [B, N, K] with per_video."""
    B, slot = dz.shape[0], dz.shape[1]
    assert dw.stride(-1) == 1 and dw.dtype == torch.float32
    a_slot = 0 if (a.shape[0] == 1 and B > 1) else a.stride(0) // _row_ld(a)
    lib = L.load()
    if (tc and pos is None and a_slot > 0 and dz.data_ptr() % 16 == 0 and a.data_ptr() % 16 == 0
            and lib.factk_wgrad_tc_supported(L.dt(dz), _row_ld(dz), L.dt(a), _row_ld(a), N, K, slot)):
        ws = _ws(dz.device, lib.factk_wgrad_tc_ws_floats(B, slot, N, K))
        COUNTERS['launches'] += 2
        _call('factk_wgrad_tc', 'wgrad_tc', dz.data_ptr(), _row_ld(dz), a.data_ptr(), _row_ld(a), a_slot, int(off), N, K, dw.data_ptr(),
              dw.stride(-2), dw.stride(0) if per_video else 0, float(alpha), int(accumulate), B, slot, L.ptr(len), ws.data_ptr(), L.stream())
        return
    ws = _ws(dz.device, lib.factk_wgrad_ws_floats(B, slot, N, K))
    COUNTERS['launches'] += 2
    _call('factk_wgrad', f"wgrad[{B}x{slot} N={N} K={K} {str(dz.dtype)[6:]},{str(a.dtype)[6:]}{' pos' if pos is not None else ''}{' bcastA' if a_slot == 0 else ''}]" if PROF_SHAPES else 'wgrad', dz.data_ptr(), L.dt(dz), _row_ld(dz), a.data_ptr(), L.dt(a), _row_ld(a), a_slot, int(off),
          L.ptr(pos), pos.stride(0) if pos is not None else 0, (pos_d if pos_d is not None else pos.shape[-1]) if pos is not None else 0,
          L.ptr(pos_idx), N, K, dw.data_ptr(), dw.stride(-2), dw.stride(0) if per_video else 0, float(alpha), int(accumulate),
          B, slot, L.ptr(len), ws.data_ptr(), L.stream())


def heads_mm(A, Bm, Cc, M, N, K, nhead, a_hs, b_hs, c_hs, a_kmajor=False, b_kmajor=False, len=None, len_mode=0, alpha=1.0,
             accumulate=False):
    """Head-batched C[b][h](m, n) (+)= alpha * sum_k A[b][h](m, k) Bm[b][h](n, k) (csrc/train_attn.cu): A, Bm, Cc rows tensors
    [B, rows, ld]; head h of an operand starts at column h * hs; k-major operands hold (m, k) at row k."""
    g = L.HeadsMM()
    B = Cc.shape[0]
    bs = lambda t: 0 if (t.shape[0] == 1 and B > 1) else t.stride(0)
    g.A, g.a_dtype, g.lda, g.a_bstride, g.a_hstride, g.a_kmajor = A.data_ptr(), L.dt(A), _row_ld(A), bs(A), a_hs, int(a_kmajor)
    g.Bm, g.b_dtype, g.ldb, g.b_bstride, g.b_hstride, g.b_kmajor = Bm.data_ptr(), L.dt(Bm), _row_ld(Bm), bs(Bm), b_hs, int(b_kmajor)
    g.C, g.c_dtype, g.ldc, g.c_bstride, g.c_hstride, g.accumulate = Cc.data_ptr(), L.dt(Cc), _row_ld(Cc), Cc.stride(0), c_hs, int(accumulate)
    g.M, g.N, g.K, g.batch, g.nhead, g.len_mode = M, N, K, B, nhead, len_mode
    g.len, g.alpha = L.ptr(len), float(alpha)
    n = L.load().factk_heads_mm_ws_floats(B, nhead, M, N, K)
    g.ws = _ws(Cc.device, n).data_ptr() if n else None
    COUNTERS['launches'] += 2 if n else 1
    _call('factk_heads_mm', f"heads_mm[{M}x{N}x{K}]" if PROF_SHAPES else 'heads_mm', g, L.stream())


def colsum(x, N, out, y=None, len=None, alpha=1.0, accumulate=True, per_video=False):
    """out[(b)][n] (+)= alpha * sum_rows x[b,t,n] (* y[b,t,n])."""
    B, slot = x.shape[0], x.shape[1]
    ws = _ws(x.device, L.load().factk_colsum_ws_floats(B, slot, N))
    COUNTERS['launches'] += 2
    _call('factk_colsum', None, x.data_ptr(), L.dt(x), _row_ld(x), L.ptr(y), L.dt(y) if y is not None else 0,
          _row_ld(y) if y is not None else 0, N, out.data_ptr(), out.stride(0) if per_video else 0, float(alpha), int(accumulate),
          B, slot, L.ptr(len), ws.data_ptr(), L.stream())


def ew(op, x, y, N, r=None, len=None, alpha=1.0, p=0.0, seed=0, site=0, bcast=False, seed_ptr=None, ridx=None):
    """Elementwise pass over the valid rows of y ([B, slot, ld]); see factk_rows_elementwise.  bcast: x is one [slot, ld] table."""
    B, slot = y.shape[0], y.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_rows_elementwise', None, int(op), x.data_ptr(), L.dt(x), _row_ld(x), L.ptr(r), L.dt(r) if r is not None else 0,
          _row_ld(r) if r is not None else 0, y.data_ptr(), L.dt(y), _row_ld(y), N, B, slot, L.ptr(len), float(alpha), float(p),
          int(seed), int(site), 0 if bcast else -1, L.ptr(seed_ptr), L.ptr(ridx), L.stream())


def transpose(src, dst):
    """dst[b, c, r] = src[b, r, c] (fp32, last dims contiguous)."""
    B, R, Cc = src.shape
    assert src.dtype == dst.dtype == torch.float32 and src.stride(-1) == 1 and dst.stride(-1) == 1
    COUNTERS['launches'] += 1
    _call('factk_transpose', None, src.data_ptr(), src.stride(1), src.stride(0), dst.data_ptr(), dst.stride(1), dst.stride(0), R, Cc, B,
          L.stream())


def splice_bwd(y, dy, dcl, dx, H, C, len=None):
    B, slot = y.shape[0], y.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_splice_bwd', None, y.data_ptr(), L.dt(y), _row_ld(y), L.ptr(dy), L.dt(dy) if dy is not None else 0,
          _row_ld(dy) if dy is not None else 0, L.ptr(dcl), _row_ld(dcl) if dcl is not None else 0, dx.data_ptr(), L.dt(dx), _row_ld(dx),
          H, C, B, slot, L.ptr(len), L.stream())


def row_softmax_bwd(p, dp, dl, M, len=None, accumulate=False):
    B, slot = p.shape[0], p.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_row_softmax_bwd', None, p.data_ptr(), _row_ld(p), dp.data_ptr(), _row_ld(dp), dl.data_ptr(), _row_ld(dl), M,
          int(accumulate), B, slot, L.ptr(len), L.stream())


def layernorm_bwd(x, w, b, dy, dv, dw, db, res=None, eps=1e-5, relu=False, len=None, accumulate=False, E=None):
    B, slot = x.shape[0], x.shape[1]
    E = x.shape[-1] if E is None else E
    ws = _ws(x.device, L.load().factk_layernorm_bwd_ws_floats(B, slot, E))
    COUNTERS['launches'] += 3
    _call('factk_layernorm_bwd', None, x.data_ptr(), L.dt(x), _row_ld(x), L.ptr(res), L.dt(res) if res is not None else 0,
          _row_ld(res) if res is not None else 0, w.data_ptr(), b.data_ptr(), float(eps), int(relu), dy.data_ptr(), L.dt(dy), _row_ld(dy),
          dv.data_ptr(), L.dt(dv), _row_ld(dv), int(accumulate), dw.data_ptr(), db.data_ptr(), B, slot, L.ptr(len), E, ws.data_ptr(),
          L.stream())


def l2norm_bwd(x, dy, dx, eps=1e-12, len=None):
    B, slot, E = x.shape
    COUNTERS['launches'] += 1
    _call('factk_l2norm_bwd', None, x.data_ptr(), L.dt(x), _row_ld(x), dy.data_ptr(), L.dt(dy), _row_ld(dy), dx.data_ptr(), L.dt(dx),
          _row_ld(dx), B, slot, L.ptr(len), E, float(eps), L.stream())


def col_softmax(logit, p, M, scale=1.0, len=None):
    B, slot = logit.shape[0], logit.shape[1]
    ws = _ws(logit.device, L.load().factk_col_softmax_train_ws_floats(B, slot, M))
    COUNTERS['launches'] += 3
    _call('factk_col_softmax', None, logit.data_ptr(), L.dt(logit), _row_ld(logit), p.data_ptr(), L.dt(p), _row_ld(p), M, float(scale), B, slot,
          L.ptr(len), ws.data_ptr(), L.stream())


def col_softmax_bwd(p, dp, dl, M, scale=1.0, len=None, accumulate=False):
    B, slot = p.shape[0], p.shape[1]
    ws = _ws(p.device, L.load().factk_colsum_ws_floats(B, slot, M) + B * M)
    COUNTERS['launches'] += 3
    _call('factk_col_softmax_bwd', None, p.data_ptr(), L.dt(p), _row_ld(p), dp.data_ptr(), L.dt(dp), _row_ld(dp), dl.data_ptr(), L.dt(dl), _row_ld(dl),
          M, float(scale), int(accumulate), B, slot, L.ptr(len), ws.data_ptr(), L.stream())


def segment_reduce(x, out, seg_start, seg_len, nseg, E, mean=False, accumulate=True):
    B, slot = x.shape[0], x.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_segment_reduce', None, x.data_ptr(), L.dt(x), _row_ld(x), out.data_ptr(), L.dt(out), _row_ld(out), seg_start.data_ptr(),
          seg_len.data_ptr(), nseg.data_ptr(), B, slot, E, int(mean), int(accumulate), L.stream())


def segment_expand(seg, seg_label, seg_len, out, E, len=None, inv_len=False, accumulate=True):
    B, slot = out.shape[0], out.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_segment_expand', None, seg.data_ptr(), L.dt(seg), _row_ld(seg), seg_label.data_ptr(), L.ptr(seg_len), out.data_ptr(),
          L.dt(out), _row_ld(out), B, slot, L.ptr(len), E, int(inv_len), int(accumulate), L.stream())


def gru_bwd(gi, gh, hout, dout, w_hh_f, w_hh_b, dgi, dgh, nseg):
    B, slot = gi.shape[0], gi.shape[1]
    Hh = w_hh_f.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_gru_bwd', None, gi.data_ptr(), gh.data_ptr(), hout.data_ptr(), L.dt(hout), _row_ld(hout), dout.data_ptr(), L.dt(dout),
          _row_ld(dout), w_hh_f.data_ptr(), w_hh_b.data_ptr(), Hh, dgi.data_ptr(), dgh.data_ptr(), B, slot, nseg.data_ptr(), L.stream())


def loss_grad_ce_rows(x, C, dx, label, cweight, nrows, coef, seg=None):
    B, slot = label.shape
    COUNTERS['launches'] += 1
    _call('factk_loss_grad_ce_rows', None, x.data_ptr(), _row_ld(x), C, dx.data_ptr(), _row_ld(dx), label.data_ptr(), cweight.data_ptr(),
          L.ptr(seg['seg_start']) if seg else None, L.ptr(seg['seg_len']) if seg else None, nrows.data_ptr(), coef.data_ptr(), B, slot,
          L.stream())


def loss_grad_smooth(x, C, dx, len, coef, mult):
    B, slot = x.shape[0], x.shape[1]
    COUNTERS['launches'] += 1
    _call('factk_loss_grad_smooth', None, x.data_ptr(), _row_ld(x), C, dx.data_ptr(), _row_ld(dx), len.data_ptr(), coef.data_ptr(),
          float(mult), B, slot, L.stream())


def loss_grad_token(aclogit, dx, aind, sind, nmatch, transcript, cweight, coef):
    B, M, C1 = aclogit.shape
    assert aclogit.is_contiguous() and dx.is_contiguous()
    COUNTERS['launches'] += 1
    _call('factk_loss_grad_token', None, aclogit.data_ptr(), M, C1, dx.data_ptr(), aind.data_ptr(), sind.data_ptr(), nmatch.data_ptr(),
          aind.shape[1], transcript.data_ptr(), transcript.shape[1], cweight.data_ptr(), coef.data_ptr(), B, L.stream())


def loss_grad_xattn(mode, x, M, dx, c, colmap, wmap, smax, nrows, coef, seg=None, mult=None, col_lse=None):
    """c: the criterion context of LossRunner.run (ground-truth segmentation of the labels)."""
    B, slot = c['gseg'].shape
    colmass = torch.empty(B, M, dtype=torch.float32, device=x.device) if mode == 1 else None
    COUNTERS['launches'] += 2 if mode == 1 else 1
    _call('factk_loss_grad_xattn', None, int(mode), x.data_ptr(), _row_ld(x), M, dx.data_ptr(), _row_ld(dx), c['gseg'].data_ptr(),
          c['gstart'].data_ptr(), c['glen'].data_ptr(), c['gn'].data_ptr(), colmap.data_ptr(), wmap.data_ptr(), smax, L.ptr(mult),
          L.ptr(colmass), L.ptr(col_lse), col_lse.stride(0) if col_lse is not None else 0,
          L.ptr(seg['seg_label']) if seg else None, L.ptr(seg['seg_start']) if seg else None, L.ptr(seg['seg_len']) if seg else None,
          nrows.data_ptr(), coef.data_ptr(), B, slot, L.stream())


def loss_grad_infonce(sim, C, ds, label, cmap, col_lse, nvalid, nseen, len, coef):
    B, slot = label.shape
    count = torch.empty(B, C, dtype=torch.float32, device=sim.device)
    COUNTERS['launches'] += 2
    _call('factk_loss_grad_infonce', None, sim.data_ptr(), _row_ld(sim), C, ds.data_ptr(), _row_ld(ds), label.data_ptr(), cmap.data_ptr(),
          count.data_ptr(), col_lse.data_ptr(), col_lse.stride(0), nvalid.data_ptr(), nseen, len.data_ptr(), coef.data_ptr(), B, slot,
          L.stream())
