"""Per-kernel parity (GPU, through the C ABI) against plain torch-CPU fp32 restatements / the oracle."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, GOLDEN

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import fact_oracle as O  # noqa: E402
from fact_clip_b200 import ops  # noqa: E402
from fact_clip_b200.ops import S  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def close(a, b, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(a.float().cpu(), b.float().cpu(), rtol=rtol, atol=atol)


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


# ------------------------------------------------------------------------------------------ gemm
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('B,slot,K,N,lens', [(1, 128, 64, 128, [128]), (3, 256, 437, 75, [256, 130, 1]),
                                              (2, 384, 2048, 256, [300, 384]), (2, 130, 20, 9, [7, 130]), (1, 64, 30, 17, [64])])
def test_gemm_plain(dtype, B, slot, K, N, lens):
    x = rnd(B, slot, K + (3 if K == 30 else 4 if K % 2 == 0 else 3), seed=1).to(dtype)   # vector and scalar load paths
    w, bias = rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    ln = torch.tensor(lens, dtype=torch.int32)
    out = torch.full((B, slot, N), 7.0, dtype=dtype, device=DEV)
    ops.gemm([S(x.to(DEV), w.to(DEV), K=K)], N, out, len=ln.to(DEV), bias=bias.to(DEV), relu=True)
    ref = torch.relu(x[..., :K].float() @ w.t() + bias)
    tol = dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    for b, T in enumerate(lens):
        close(out[b, :T], ref[b, :T], **tol)
        assert bool((out[b, T:].float() == 7.0).all())          # rows >= len are never written


def test_gemm_conv_taps_and_residual():
    """Three shifted taps == nn.Conv1d(k=3, dilation=d, padding=d); then 1x1 + residual (basic.py:158-164)."""
    B, slot, F, lens = 2, 256, 32, [256, 100]
    x = rnd(B, slot, F, seed=4)
    conv = torch.nn.Conv1d(F, F, 3, padding=8, dilation=8)
    w3 = conv.weight.detach().permute(2, 0, 1).contiguous()
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    xd = x.to(DEV)
    h = torch.zeros(B, slot, F, device=DEV)
    ops.gemm([S(xd, w3[k].to(DEV), off=(k - 1) * 8) for k in range(3)], F, h, len=ln, bias=conv.bias.detach().to(DEV), relu=True)
    w1 = rnd(F, F, seed=5, scale=F ** -0.5)
    y = torch.zeros(B, slot, F, device=DEV)
    ops.gemm([S(h, w1.to(DEV))], F, y, len=ln, res=xd)
    for b, T in enumerate(lens):
        hr = torch.relu(conv(x[b, :T].t()[None])[0].t()).detach()
        close(h[b, :T], hr)
        close(y[b, :T], hr @ w1.t() + x[b, :T])


def test_gemm_gather_pos_pervideo_weights():
    B, slot, K1, K2, N, M = 2, 128, 16, 24, 20, 10
    lens = [128, 50]
    seg = rnd(B, slot, K1, seed=6)
    fr = rnd(B, slot, K2, seed=7)
    idx = torch.randint(0, 40, (B, slot), generator=torch.Generator().manual_seed(8), dtype=torch.int32)
    pos = rnd(slot, K2, seed=9)
    pidx = torch.randint(0, slot, (B, slot), generator=torch.Generator().manual_seed(10), dtype=torch.int32)
    w1 = rnd(N, K1, seed=11)
    w2 = rnd(B, N, K2, seed=12)             # per-video weights
    bias = rnd(B, N, seed=13)
    out = torch.zeros(B, slot, N, device=DEV)
    ops.gemm([S(seg.to(DEV), w1.to(DEV), gather=idx.to(DEV)),
              S(fr.to(DEV), w2.to(DEV), pos=pos.to(DEV), pos_d=12, pos_idx=pidx.to(DEV))], N, out,
             len=torch.tensor(lens, dtype=torch.int32, device=DEV), bias=bias.to(DEV), alpha=0.5)
    for b, T in enumerate(lens):
        a2 = fr[b, :T].clone()
        a2[:, :12] += pos[pidx[b, :T].long()][:, :12]
        ref = 0.5 * (seg[b][idx[b, :T].long()] @ w1.t() + a2 @ w2[b].t()) + bias[b]
        close(out[b, :T], ref)


def test_gemm_shared_A_rows():
    """A with a single 'video' broadcast against per-video weights (used for vt = Wa xv^T)."""
    B, Fd, H, M = 3, 40, 32, 11
    Wa, xv = rnd(1, Fd, H, seed=14), rnd(B, M, H, seed=15)
    out = torch.zeros(B, Fd, 12, device=DEV)
    ops.gemm([S(Wa.to(DEV), xv.to(DEV))], M, out)
    close(out[:, :, :M], torch.einsum('fh,bmh->bfm', Wa[0], xv))


# ------------------------------------------------------------------------------------------ row ops
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_softmax_splice(dtype):
    B, slot, H, C, lens = 2, 64, 48, 11, [64, 9]
    x = rnd(B, slot, H, seed=16, scale=2.0).to(dtype)
    xd = x.to(DEV).clone()
    clogit = torch.zeros(B, slot, C, device=DEV)
    pred = torch.full((B, slot), -1, dtype=torch.int32, device=DEV)
    ops.softmax_splice(xd, C, clogit, pred, len=torch.tensor(lens, dtype=torch.int32, device=DEV))
    for b, T in enumerate(lens):
        feat, cl = O.process_feature(x[b, :T].float(), C)
        close(clogit[b, :T], cl, rtol=0, atol=0)
        close(xd[b, :T], feat, rtol=1e-2 if dtype == torch.bfloat16 else 1e-5, atol=1e-2 if dtype == torch.bfloat16 else 1e-6)
        assert torch.equal(pred[b, :T].cpu().long(), cl.argmax(-1))
        close(xd[b, T:], x[b, T:], rtol=0, atol=0)


def test_softmax_splice_frame_level_four_rows_per_warp():
    """The frame-level launches (bf16, 65..96 classes, >= 32768 rows) take splice_multi_kernel (four rows per warp): same results as
    the oracle's process_feature, ragged lengths that end inside a four-row group, rows past the end untouched."""
    B, slot, H, C, lens = 3, 16384, 512, 75, [16384, 16381, 5]
    x = rnd(B, slot, H, seed=61, scale=2.0).to(torch.bfloat16)
    xd = x.to(DEV).clone()
    clogit = torch.zeros(B, slot, C, device=DEV)
    pred = torch.full((B, slot), -1, dtype=torch.int32, device=DEV)
    ops.softmax_splice(xd, C, clogit, pred, len=torch.tensor(lens, dtype=torch.int32, device=DEV))
    for b, T in enumerate(lens):
        feat, cl = O.process_feature(x[b, :T].float(), C)
        close(clogit[b, :T], cl, rtol=0, atol=0)
        close(xd[b, :T], feat, rtol=1e-2, atol=1e-2)
        assert torch.equal(pred[b, :T].cpu().long(), cl.argmax(-1))
        close(xd[b, T:], x[b, T:], rtol=0, atol=0)
        assert bool((pred[b, T:] == -1).all())


def test_layernorm_l2norm_rowsoftmax_gather():
    B, slot, E = 2, 40, 96
    x, r = rnd(B, slot, E, seed=17), rnd(B, slot, E, seed=18)
    w, bb = rnd(E, seed=19), rnd(E, seed=20)
    y = torch.zeros(B, slot, E, device=DEV)
    ops.layernorm(x.to(DEV), w.to(DEV), bb.to(DEV), y, res=r.to(DEV), relu=True)
    close(y, torch.relu(torch.nn.functional.layer_norm(x + r, (E,), w, bb)), rtol=1e-4, atol=1e-5)
    ops.l2norm(x.to(DEV), y)
    close(y, torch.nn.functional.normalize(x, dim=-1))
    p = torch.zeros(B, slot, E, device=DEV)
    ops.row_softmax(x.to(DEV), p, 75, scale=0.5)
    close(p[..., :75], torch.softmax(0.5 * x[..., :75], -1))
    idx = torch.randint(0, slot, (B, slot), generator=torch.Generator().manual_seed(21), dtype=torch.int32)
    g = torch.zeros(B, slot, E, device=DEV)
    ops.gather_rows(x.to(DEV), idx.to(DEV), g, E)
    close(g, torch.stack([x[b][idx[b].long()] for b in range(B)]), rtol=0, atol=0)


# ------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize('M,nh,dh', [(75, 8, 32), (12, 4, 8), (300, 8, 32), (60, 8, 64)])
def test_mha_tokens(M, nh, dh):
    B, E = 2, nh * dh
    qkv = rnd(B, M, 3 * E, seed=22)
    d = qkv.to(DEV)
    o = torch.zeros(B, M, E, device=DEV)
    ops.mha_tokens(d[..., :E], d[..., E:2 * E], d[..., 2 * E:], o, nh)
    q, k, v = [t.view(B, M, nh, dh).transpose(1, 2) for t in (qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:])]
    ref = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), -1) @ v).transpose(1, 2).reshape(B, M, E)
    close(o, ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('M,nh,dh', [(75, 8, 32), (1, 8, 32), (300, 8, 32), (60, 8, 64), (83, 4, 16)])
def test_mha_tokens_tf32_tensor_core(M, nh, dh):
    """tf32 mma.sync variant (bf16 compute mode): tf32 operands, fp32 accumulate -> 2e-3 relative to the fp32 reference."""
    B, E = 3, nh * dh
    qkv = rnd(B, M, 3 * E, seed=23)
    d = qkv.to(DEV)
    o = torch.zeros(B, M, E, device=DEV)
    ops.mha_tokens(d[..., :E], d[..., E:2 * E], d[..., 2 * E:], o, nh, tf32=True)
    q, k, v = [t.view(B, M, nh, dh).transpose(1, 2) for t in (qkv[..., :E], qkv[..., E:2 * E], qkv[..., 2 * E:])]
    ref = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), -1) @ v).transpose(1, 2).reshape(B, M, E)
    err = float((o.cpu() - ref).norm() / ref.norm())
    assert err < 2e-3, err


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,nh,dh,slot,lens', [(75, 8, 32, 1280, [1280, 700, 3]), (12, 4, 8, 128, [37, 128]), (300, 8, 32, 640, [640])])
def test_attn_rows(dtype, M, nh, dh, slot, lens):
    B, E = len(lens), nh * dh
    q = rnd(B, M, E, seed=23)
    kv = rnd(B, slot, 2 * E, seed=24).to(dtype)
    o = torch.zeros(B, M, E, device=DEV)
    kvd = kv.to(DEV)
    ws = torch.empty(ops.attn_rows_ws(B, slot, M, nh, dh), device=DEV)
    ops.attn_rows(q.to(DEV), kvd[..., :E], kvd[..., E:], o, nh, ws, len=torch.tensor(lens, dtype=torch.int32, device=DEV))
    for b, T in enumerate(lens):
        qq = q[b].view(M, nh, dh).transpose(0, 1)
        kk = kv[b, :T, :E].float().view(T, nh, dh).transpose(0, 1)
        vv = kv[b, :T, E:].float().view(T, nh, dh).transpose(0, 1)
        ref = (torch.softmax(qq @ kk.transpose(1, 2) / math.sqrt(dh), -1) @ vv).transpose(0, 1).reshape(M, E)
        # bf16 keys/values take the HMMA kernel: probabilities are rounded to bf16 before the PV product
        if dtype == torch.float32:
            close(o[b], ref, rtol=2e-4, atol=2e-5)
        else:
            assert rel_l2(o[b], ref) < 5e-3


@pytest.mark.parametrize('M,slot,lens', [(75, 1280, [1280, 700, 3]), (128, 512, [512, 65]), (300, 640, [640]), (129, 1024, [1000, 513])])
def test_attn_rows_tcgen05_kernel(M, slot, lens, monkeypatch):
    """The tcgen05 / TMEM attention kernel (attn_tc.cu) forced on for every M (it is the default only for M > 128): S = Q K^T
    and O += P V on the tensor cores, softmax on TMEM rows in registers -- against fp32 softmax attention in torch."""
    import subprocess, sys, os
    code = f"""
import math, sys, torch
sys.path.insert(0, {ROOT!r})
from fact_clip_b200 import ops
M, slot, lens, nh, dh = {M}, {slot}, {lens}, 8, 32
B, E = len(lens), nh * dh
g = torch.Generator().manual_seed(23)
q = torch.randn(B, M, E, generator=g)
kv = torch.randn(B, slot, 2 * E, generator=g).to(torch.bfloat16)
o = torch.zeros(B, M, E, device='cuda')
kvd = kv.cuda()
ws = torch.empty(ops.attn_rows_ws(B, slot, M, nh, dh), device='cuda')
ops.attn_rows(q.cuda(), kvd[..., :E], kvd[..., E:], o, nh, ws, len=torch.tensor(lens, dtype=torch.int32, device='cuda'))
for b, T in enumerate(lens):
    qq = q[b].view(M, nh, dh).transpose(0, 1)
    kk = kv[b, :T, :E].float().view(T, nh, dh).transpose(0, 1)
    vv = kv[b, :T, E:].float().view(T, nh, dh).transpose(0, 1)
    ref = (torch.softmax(qq @ kk.transpose(1, 2) / math.sqrt(dh), -1) @ vv).transpose(0, 1).reshape(M, E)
    err = float((o[b].cpu() - ref).norm() / ref.norm())
    assert err < 5e-3, (b, err)
print('ok')
"""
    p = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, FACTK_ATTN_TC='1'), capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and 'ok' in p.stdout, p.stderr[-2000:]


@pytest.mark.parametrize('M,E,slot,lens', [(75, 512, 1280, [1280, 600, 1]), (12, 64, 128, [100, 128])])
def test_col_softmax_apply(M, E, slot, lens):
    B, Mp = len(lens), (M + 3) // 4 * 4
    logit = rnd(B, slot, Mp, seed=25, scale=3.0)
    x = rnd(B, slot, E, seed=26)
    out = torch.zeros(B, M, E, device=DEV)
    attn = torch.zeros(B, slot, Mp, device=DEV)
    ws = torch.empty(ops.col_softmax_ws(B, slot, M, E), device=DEV)
    ops.col_softmax_apply(logit.to(DEV), x.to(DEV), out, M, ws, attn=attn, len=torch.tensor(lens, dtype=torch.int32, device=DEV))
    for b, T in enumerate(lens):
        p = torch.softmax(logit[b, :T, :M], dim=0)
        close(attn[b, :T, :M], p, rtol=1e-4, atol=1e-6)
        close(out[b], p.t() @ x[b, :T], rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize('M,E,slot,lens', [(75, 512, 1280, [1280, 600, 1]), (300, 512, 640, [640, 77]), (12, 64, 128, [100, 128])])
def test_col_softmax_apply_bf16_rows(M, E, slot, lens):
    """bf16 rows take the mma.sync kernel (probabilities rounded to bf16 before the product); garbage past the end is ignored."""
    B, Mp = len(lens), (M + 3) // 4 * 4
    logit = rnd(B, slot, Mp, seed=25, scale=3.0)
    x = rnd(B, slot, E, seed=26).to(torch.bfloat16)
    xd = x.to(DEV).clone()
    for b, T in enumerate(lens):
        xd[b, T:] = float('nan')            # rows past len[b] may hold anything
    out = torch.zeros(B, M, E, device=DEV)
    attn = torch.zeros(B, slot, Mp, device=DEV)
    ws = torch.empty(ops.col_softmax_ws(B, slot, M, E), device=DEV)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    ops.col_softmax_apply(logit.to(DEV), xd, out, M, ws, attn=attn, len=ln)
    out2 = torch.zeros(B, M, E, device=DEV)
    ops.col_softmax_apply(logit.to(DEV), xd, out2, M, ws, len=ln)
    assert torch.equal(out, out2)
    for b, T in enumerate(lens):
        p = torch.softmax(logit[b, :T, :M], dim=0)
        close(attn[b, :T, :M], p, rtol=1e-4, atol=1e-6)
        ref = p.t() @ x[b, :T].float()
        assert rel_l2(out[b], ref) < 4e-3, rel_l2(out[b], ref)


# ------------------------------------------------------------------------------------------ TDU
def test_tdu_segment_and_mean():
    g = torch.Generator().manual_seed(27)
    B, slot, E = 4, 2304, 64
    lens = [2304, 1500, 1, 1025]
    pred = torch.zeros(B, slot, dtype=torch.int32)
    for b in range(B):
        # run lengths 1..7 -> many segments; video 2 has a single frame; video 3 one long run
        runs = torch.randint(1, 8, (slot,), generator=g)
        cls = torch.randint(0, 5, (slot,), generator=g)
        p = torch.repeat_interleave(cls, runs)[:slot]
        pred[b] = p.int() if b != 3 else 2
    x = rnd(B, slot, E, seed=28)
    names = ['seg_label', 'seg_start', 'seg_len', 'seg_center']
    o = {n: torch.full((B, slot), -1, dtype=torch.int32, device=DEV) for n in names}
    nseg = torch.zeros(B, dtype=torch.int32, device=DEV)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    ops.tdu_segment(pred.to(DEV), o['seg_label'], o['seg_start'], o['seg_len'], o['seg_center'], nseg, len=ln)
    seg = torch.zeros(B, slot, E, device=DEV)
    ops.segment_mean(x.to(DEV), seg, o['seg_label'], o['seg_start'], o['seg_len'], nseg)
    # streaming two-pass kernel (workspace given), fp32 and bf16 rows
    ws = torch.empty(ops.segment_mean_ws(B, slot, E), device=DEV)
    seg_s = torch.zeros(B, slot, E, device=DEV)
    ops.segment_mean(x.to(DEV), seg_s, o['seg_label'], o['seg_start'], o['seg_len'], nseg, ws=ws)
    seg_s2 = torch.zeros(B, slot, E, device=DEV)
    ops.segment_mean(x.to(DEV), seg_s2, o['seg_label'], o['seg_start'], o['seg_len'], nseg, ws=ws)
    assert torch.equal(seg_s, seg_s2)
    x16 = x.to(torch.bfloat16)
    seg_h = torch.zeros(B, slot, E, device=DEV, dtype=torch.bfloat16)
    ops.segment_mean(x16.to(DEV), seg_h, o['seg_label'], o['seg_start'], o['seg_len'], nseg, ws=ws)
    for b, T in enumerate(lens):
        lab, start, slen = O.run_length(pred[b, :T].numpy())
        S = len(slen)
        assert int(nseg[b]) == S
        assert np.array_equal(o['seg_label'][b, :T].cpu().numpy(), lab)
        assert np.array_equal(o['seg_start'][b, :S].cpu().numpy(), start)
        assert np.array_equal(o['seg_len'][b, :S].cpu().numpy(), slen)
        assert np.array_equal(o['seg_center'][b, :S].cpu().numpy(), (start + start + slen - 1) // 2)
        ref = torch.zeros(S, E).index_add_(0, torch.from_numpy(lab), x[b, :T]) / torch.from_numpy(slen)[:, None]
        close(seg[b, :S], ref, rtol=1e-4, atol=1e-5)
        close(seg_s[b, :S], ref, rtol=1e-4, atol=1e-5)
        ref16 = torch.zeros(S, E).index_add_(0, torch.from_numpy(lab), x16[b, :T].float()) / torch.from_numpy(slen)[:, None]
        close(seg_h[b, :S].float(), ref16, rtol=1e-2, atol=1e-2)
        assert bool((seg_s[b, S:] == 0).all())
    # determinism: bit-identical on a second run
    seg2 = torch.zeros(B, slot, E, device=DEV)
    ops.segment_mean(x.to(DEV), seg2, o['seg_label'], o['seg_start'], o['seg_len'], nseg)
    assert torch.equal(seg, seg2)


@pytest.mark.parametrize('Hh,nsegs', [(32, [19, 1, 7, 40, 3]), (256, [300, 17])])
def test_gru_bidir(Hh, nsegs):
    B, slot, H = len(nsegs), 384, 2 * Hh
    gru = torch.nn.GRU(H, Hh, 1, bidirectional=True)
    sd = {'g.' + k: v.detach() for k, v in gru.state_dict().items()}
    x = rnd(B, slot, H, seed=29)
    wih = torch.cat([sd['g.weight_ih_l0'], sd['g.weight_ih_l0_reverse']])
    bih = torch.cat([sd['g.bias_ih_l0'], sd['g.bias_ih_l0_reverse']])
    gi = (x @ wih.t() + bih).to(DEV)
    out = torch.zeros(B, slot, H, device=DEV)
    ops.gru_bidir(gi, sd['g.weight_hh_l0'].to(DEV), sd['g.bias_hh_l0'].to(DEV), sd['g.weight_hh_l0_reverse'].to(DEV),
                  sd['g.bias_hh_l0_reverse'].to(DEV), out, torch.tensor(nsegs, dtype=torch.int32, device=DEV), relu=False)
    for b, Sg in enumerate(nsegs):
        ref = O.gru_bidir_fast(sd, 'g.', x[b, :Sg])
        close(out[b, :Sg], ref, rtol=2e-4, atol=2e-5)
        assert bool((out[b, Sg:] == 0).all())


@pytest.mark.parametrize('nsegs', [[300, 17], [5, 1, 64, 33, 200, 7, 90, 2, 11, 384],
                                   [(7 * i) % 41 + 1 for i in range(83)]])   # > 72 videos: two interleaved chains per cluster
def test_gru_bidir_mma(nsegs):
    """Tensor-core GRU (bf16 W_hh and exchanged state, MUFU.TANH gates) vs the fp32 recurrence: bf16-level agreement."""
    Hh = 256
    B, slot, H = len(nsegs), 384, 2 * Hh
    gru = torch.nn.GRU(H, Hh, 1, bidirectional=True)
    sd = {'g.' + k: v.detach() for k, v in gru.state_dict().items()}
    x = rnd(B, slot, H, seed=31)
    wih = torch.cat([sd['g.weight_ih_l0'], sd['g.weight_ih_l0_reverse']])
    bih = torch.cat([sd['g.bias_ih_l0'], sd['g.bias_ih_l0_reverse']])
    gi = (x @ wih.t() + bih).to(DEV)
    ns = torch.tensor(nsegs, dtype=torch.int32, device=DEV)
    args = (gi, sd['g.weight_hh_l0'].to(DEV), sd['g.bias_hh_l0'].to(DEV), sd['g.weight_hh_l0_reverse'].to(DEV),
            sd['g.bias_hh_l0_reverse'].to(DEV))
    out = torch.zeros(B, slot, H, device=DEV)
    ops.gru_bidir(*args, out, ns, relu=False, mma=True)
    out16 = torch.zeros(B, slot, H, device=DEV, dtype=torch.bfloat16)
    ops.gru_bidir(*args, out16, ns, relu=True, mma=True)
    for b, Sg in enumerate(nsegs):
        ref = O.gru_bidir_fast(sd, 'g.', x[b, :Sg])
        err = float((out[b, :Sg].cpu() - ref).abs().max())
        assert err < 1e-2, (b, Sg, err)
        assert float((out[b, :Sg].cpu() - ref).norm() / ref.norm()) < 5e-3
        close(out16[b, :Sg].float(), torch.relu(ref), rtol=1e-2, atol=1e-2)
        assert bool((out[b, Sg:] == 0).all())
    out2 = torch.zeros(B, slot, H, device=DEV)
    ops.gru_bidir(*args, out2, ns, relu=False, mma=True)
    assert torch.equal(out, out2)
    # videos grouped by decreasing segment count (ranked on the device): every video's chain is computed in its own MMA
    # column, so the grouping must not change a single bit
    out3 = torch.zeros(B, slot, H, device=DEV)
    order = torch.full((B,), -1, dtype=torch.int32, device=DEV)
    ops.gru_bidir(*args, out3, ns, relu=False, mma=True, order_ws=order)
    assert torch.equal(out, out3)
    rank = sorted(range(B), key=lambda i: (-nsegs[i], i))
    assert order.tolist() == rank


# ------------------------------------------------------------------------------------------ eval
def test_fuse_eval_known_answers():
    cases = torch.load(os.path.join(GOLDEN, 'eval_cases.pt'), weights_only=False)
    for c in cases:
        ac, attn, fc = c['action_clogit'][:, 0], c['a2f_attn'][0], c['frame_clogit'][:, 0]
        T, M, Cc = fc.shape[0], ac.shape[0], fc.shape[1]
        pred = torch.full((1, T), -1, dtype=torch.int64, device=DEV)
        ops.fuse_eval(ac[None].contiguous().to(DEV), attn[None].contiguous().to(DEV), fc[None].contiguous().to(DEV),
                      c['weight'], pred, M, Cc)
        assert torch.equal(pred[0].cpu(), c['pred'])


def test_fuse_eval_segment_gather():
    g = torch.Generator().manual_seed(30)
    T, Sg, M, Cc = 50, 9, 7, 5
    ac = torch.randn(M, Cc + 1, generator=g) * 3
    attn_seg = torch.softmax(torch.randn(Sg, M, generator=g) * 2, -1)
    lab = torch.sort(torch.randint(0, Sg, (T,), generator=g)).values
    fl = torch.randn(T, Cc, generator=g)
    pred = torch.zeros(1, T, dtype=torch.int64, device=DEV)
    ops.fuse_eval(ac[None].to(DEV), attn_seg[None].contiguous().to(DEV), fl[None].to(DEV), 0.1, pred, M, Cc,
                  seg_label=lab.int()[None].to(DEV))
    ref = O.fuse_eval(ac, attn_seg[lab], torch.softmax(fl, -1), 0.1)
    assert torch.equal(pred[0].cpu(), ref)


def test_fuse_eval_transcript_and_embed_tokens():
    """FACT.trans pieces: _eval_w_transcript (blocks.py:263-275) frame-level and through a segment gather, and the
    token initialisation action_embed(transcript) + action_pe (blocks.py:75-78)."""
    g = torch.Generator().manual_seed(31)
    Cc, A = 9, 32
    for T, N, Sg in ((70, 6, 11), (3, 1, 2), (200, 25, 40)):
        tr = torch.randint(0, Cc, (N,), generator=g)
        fl = torch.randn(T, Cc, generator=g) * 2
        attn = torch.softmax(torch.randn(T, N, generator=g) * 2, -1)
        ntr = torch.tensor([N], dtype=torch.int32, device=DEV)
        pred = torch.full((1, T), -1, dtype=torch.int64, device=DEV)
        ops.fuse_eval_transcript(attn[None].contiguous().to(DEV), fl[None].to(DEV), 0.1, tr.int()[None].to(DEV), ntr, pred, Cc)
        assert torch.equal(pred[0].cpu(), O.eval_w_transcript(tr, attn, fl, 0.1))
        attn_seg = torch.softmax(torch.randn(Sg, N, generator=g) * 2, -1)
        lab = torch.sort(torch.randint(0, Sg, (T,), generator=g)).values
        ops.fuse_eval_transcript(attn_seg[None].contiguous().to(DEV), fl[None].to(DEV), 0.1, tr.int()[None].to(DEV), ntr, pred, Cc,
                                 seg_label=lab.int()[None].to(DEV))
        assert torch.equal(pred[0].cpu(), O.eval_w_transcript(tr, attn_seg[lab], fl, 0.1))
        emb = torch.randn(Cc, A, generator=g)
        pe = O.positional_table(A, 64)
        out = torch.empty(1, N, A, device=DEV)
        ops.embed_tokens(emb.to(DEV), tr.int().to(DEV), pe.to(DEV), out)
        assert torch.equal(out[0].cpu(), emb[tr] + pe[:N])


@pytest.mark.parametrize('C,M,f_logp', [(300, 40, False), (300, 40, True), (97, 12, True), (1200, 75, True)])
def test_fuse_eval_wide_class_rows(C, M, f_logp):
    """Prob fusion with more classes than the register-resident kernel holds (the Epic action table has thousands): generic
    kernel, with softmaxed frame logits and with un-normalised frame log-probabilities (blocks_SepVerbNoun.py:307-329)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import vn_oracle as VO
    g = torch.Generator().manual_seed(41)
    T = 150
    ac = torch.randn(M, C + 1, generator=g) * 3
    ac[::3, -1] += 6.0                                            # some tokens predict the null class
    attn = torch.softmax(torch.randn(T, M, generator=g) * 2, -1)
    fl = torch.randn(T, C, generator=g) * 2
    pred = torch.full((1, T), -1, dtype=torch.int64, device=DEV)
    if f_logp:
        alp, flp = torch.log_softmax(ac, -1), torch.log_softmax(fl, -1) - 0.3       # frame rows do NOT sum to one
        ops.fuse_eval(alp[None].to(DEV), attn[None].contiguous().to(DEV), flp[None].to(DEV), 0.1, pred, M, C, f_logp=True)
        ref = VO.evaluate(alp, attn, flp, 0.1)
    else:
        ops.fuse_eval(ac[None].to(DEV), attn[None].contiguous().to(DEV), fl[None].to(DEV), 0.1, pred, M, C)
        ref = O.fuse_eval(ac, attn, torch.softmax(fl, -1), 0.1)
    agree = float((pred[0].cpu() == ref).float().mean())
    assert agree >= 0.99, agree                                   # __expf vs exp can flip a near-tie among hundreds of classes


@pytest.mark.parametrize('M,slot,lens,with_outputs', [(75, 512, [512, 300], True), (128, 256, [256], True), (12, 384, [1, 384, 130], False),
                                                       (64, 128, [100], True), (75, 4096, [4096, 4000], False)])
def test_a2f_fused(M, slot, lens, with_outputs):
    """Fused X2Y_map, a2f direction (x2y_fused.cu): logits -> row softmax -> value GEMM + Y_W in one tcgen05 kernel, against fp32 torch."""
    B, H, F = len(lens), 512, 256
    Kp, Mp = (M + 63) // 64 * 64, (M + 3) // 4 * 4
    bf = torch.bfloat16
    rows = rnd(B, slot, H, seed=31).to(bf)
    kt = (rnd(B, M, H, seed=32) * 0.1).to(bf)
    cb = rnd(B, M, seed=33)
    wy = (rnd(F, H, seed=34) * H ** -0.5).to(bf)
    vt = torch.zeros(B, F, Kp, dtype=bf)
    vt[:, :, :M] = rnd(B, F, M, seed=35).to(bf)
    bias = rnd(F, seed=36)
    out = torch.zeros(B, slot, F, dtype=bf, device=DEV)
    logit = torch.zeros(B, slot, Mp, device=DEV) if with_outputs else None
    attn = torch.zeros(B, slot, Mp, device=DEV) if with_outputs else None
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    ops.a2f_fused(rows.to(DEV), kt.to(DEV), cb.to(DEV), wy.to(DEV), vt.to(DEV), bias.to(DEV), out, M, logit=logit, attn=attn, len=ln)
    for b, T in enumerate(lens):
        x = rows[b, :T].float()
        lg = x @ kt[b].float().t() + cb[b]
        p = torch.softmax(lg, -1)
        ref = x @ wy.float().t() + p.to(bf).float() @ vt[b, :, :M].float().t() + bias
        assert rel_l2(out[b, :T].float(), ref) < 6e-3, (b, rel_l2(out[b, :T].float(), ref))
        if with_outputs:
            assert rel_l2(logit[b, :T, :M], lg) < 1e-5 and rel_l2(attn[b, :T, :M], p) < 1e-4
        assert float(out[b, T:].float().abs().sum()) == 0.0          # rows past the end are never written


@pytest.mark.parametrize('M,H,slot,lens,qscale,ramp', [
    (75, 512, 4096, [4096, 4000, 577], 0.05, 0.0), (75, 512, 512, [512, 300], 0.1, 0.0), (80, 512, 256, [256, 1, 129], 0.1, 0.0),
    (12, 512, 384, [64, 384, 130], 0.3, 0.0), (33, 512, 640, [640, 65], 0.1, -25.0), (75, 512, 2304, [2304, 1500], 0.1, 40.0),
    (75, 256, 512, [512, 300], 0.1, 0.0), (128, 256, 1024, [1024, 513, 1], 0.1, 0.0), (96, 256, 2048, [2048, 1500], 0.1, 40.0),
    (12, 256, 384, [64, 384, 130], 0.3, 0.0), (33, 256, 640, [640, 65], 0.1, -25.0)])
def test_f2a_fused(M, H, slot, lens, qscale, ramp):
    """Fused X2Y_map, f2a direction (f2a_fused_t.cu: rows on the M side, hid_dim 512 / 256; f2a_fused.cu: tokens on the M side,
    hid_dim 256 with more than 80 tokens): S = qt rows^T -> softmax over the rows (online, 2048-row CTAs) -> weighted row sum in
    one tcgen05 kernel + the split combine, against fp64 torch on the same bf16 operands.  ``ramp`` adds a trend along the rows
    so that the running maximum keeps growing (positive: the accumulator rescale path) or the first tile dominates (negative);
    rows past len[b] hold NaN (they must not reach the accumulator)."""
    B = len(lens)
    bf = torch.bfloat16
    assert ops.f2a_fused_ok(M, H, slot)
    rows = rnd(B, slot, H, seed=41)
    qt = rnd(B, M, H, seed=42) * qscale * (256 / H) ** 0.5
    if ramp:
        # one channel carries t / slot; the queries read it with weight ``ramp``: logit += ramp * t / slot
        rows[:, :, 0] = torch.arange(slot, dtype=torch.float32)[None, :] / slot
        qt[:, :, 0] = ramp
    rows, qt = rows.to(bf), qt.to(bf)
    rows_dev = rows.clone()
    for b, T in enumerate(lens):
        rows_dev[b, T:] = float('nan')
    out = torch.full((B, M, H), float('nan'), device=DEV)
    ws = torch.empty(ops.f2a_fused_ws(B, slot, M, H), device=DEV)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    ops.f2a_fused(rows_dev.to(DEV), qt.to(DEV), out, M, ws, len=ln)
    torch.cuda.synchronize()
    for b, T in enumerate(lens):
        x = rows[b, :T].double()
        p = torch.softmax(x @ qt[b].double().t(), 0)                # [T, M], softmax over the rows
        ref = (p.t() @ x).float()
        err = rel_l2(out[b], ref)
        assert err < 6e-3, (b, T, err)
    # bit-reproducible and batch invariant: video 0 alone gives the same bits
    out1 = torch.empty(1, M, H, device=DEV)
    ops.f2a_fused(rows_dev[:1].contiguous().to(DEV), qt[:1].contiguous().to(DEV), out1, M, ws, len=ln[:1].contiguous())
    assert torch.equal(out1[0], out[0])


def _ref_token_layer(x, nhead, Wo, bo, ln1, Win=None, bin_=None, pos=None, o_in=None, Wq=None, bq=None, ffn=None):
    """fp32 torch restatement of SALayer / the two halves of SCALayer (models/basic.py:429-452, 494-523), eval mode."""
    import torch.nn.functional as Fn
    B, M, A = x.shape
    if Win is not None:
        xq = x + (pos if pos is not None else 0)
        q, k = (xq @ Win[:2 * A].t() + bin_[:2 * A]).split(A, -1)
        v = x @ Win[2 * A:].t() + bin_[2 * A:]
        hd = lambda t: t.view(B, M, nhead, A // nhead).transpose(1, 2)
        p = torch.softmax(hd(q) @ hd(k).transpose(-1, -2) / (A // nhead) ** 0.5, -1)
        o = (p @ hd(v)).transpose(1, 2).reshape(B, M, A)
    else:
        o = o_in
    x1 = Fn.layer_norm(o @ Wo.t() + bo + x, (A,), ln1[0], ln1[1])
    cq = None
    if Wq is not None:
        cq = (x1 + (pos if pos is not None else 0)) @ Wq.t() + bq
    if ffn is not None:
        W1, b1, W2, b2, l2w, l2b = ffn
        x1 = Fn.layer_norm(torch.relu(x1 @ W1.t() + b1) @ W2.t() + b2 + x1, (A,), l2w, l2b)
    return x1, cq


@pytest.mark.parametrize('B,M,A,nhead,ff', [(3, 75, 256, 8, 512), (2, 12, 64, 4, 64), (1, 80, 128, 2, 128), (2, 33, 256, 8, 256)])
def test_token_layer_fused(B, M, A, nhead, ff):
    """csrc/token_layer.cu: a whole SALayer, and the two halves of an SCALayer, each in ONE launch, against fp32 torch (bf16 GEMM
    operands, fp32 accumulate / residual / LayerNorm / softmax: 1e-2 relative)."""
    assert ops.token_layer_ok(M, A, nhead, ff)
    mk = lambda *s, seed, sc=1.0: (rnd(*s, seed=seed) * sc).to(DEV)
    x = mk(B, M, A, seed=1)
    pos = mk(M, A, seed=2, sc=0.5)
    Win, bin_ = mk(3 * A, A, seed=3, sc=A ** -0.5), mk(3 * A, seed=4, sc=0.1)
    Wo, bo = mk(A, A, seed=5, sc=A ** -0.5), mk(A, seed=6, sc=0.1)
    Wq, bq = mk(A, A, seed=7, sc=A ** -0.5), mk(A, seed=8, sc=0.1)
    W1, b1 = mk(ff, A, seed=9, sc=A ** -0.5), mk(ff, seed=10, sc=0.1)
    W2, b2 = mk(A, ff, seed=11, sc=ff ** -0.5), mk(A, seed=12, sc=0.1)
    ln = [(1 + mk(A, seed=20 + i, sc=0.1), mk(A, seed=30 + i, sc=0.1)) for i in range(3)]
    pk = ops.pack_token_weight
    pre_qk, pre_q = (pos @ Win[:2 * A].t()).contiguous(), (pos @ Wq.t()).contiguous()
    # SALayer
    ref, _ = _ref_token_layer(x, nhead, Wo, bo, ln[0], Win=Win, bin_=bin_, pos=pos, ffn=(W1, b1, W2, b2, *ln[1]))
    got = x.clone()
    ops.token_layer(got, nhead, pk(Wo), bo, *ln[0], w_in=pk(Win), b_in=bin_, pre_qk=pre_qk, ffn=(pk(W1), b1, pk(W2), b2, *ln[1], ff))
    assert rel_l2(got, ref) < 1e-2, rel_l2(got, ref)
    # SCALayer, first half: self attention + norm1 + query projection
    ref1, refq = _ref_token_layer(x, nhead, Wo, bo, ln[0], Win=Win, bin_=bin_, pos=pos, Wq=Wq, bq=bq)
    got1, cq = x.clone(), torch.zeros_like(x)
    ops.token_layer(got1, nhead, pk(Wo), bo, *ln[0], w_in=pk(Win), b_in=bin_, pre_qk=pre_qk, w_q=pk(Wq), b_q=bq, pre_q=pre_q, cq_out=cq)
    assert rel_l2(got1, ref1) < 1e-2 and rel_l2(cq, refq) < 1e-2, (rel_l2(got1, ref1), rel_l2(cq, refq))
    # SCALayer, second half: out_proj of the cross attention's output + norm2 + FFN + norm3
    o_in = mk(B, M, A, seed=40)
    ref2, _ = _ref_token_layer(x, nhead, Wo, bo, ln[2], o_in=o_in, ffn=(W1, b1, W2, b2, *ln[1]))
    got2 = x.clone()
    ops.token_layer(got2, nhead, pk(Wo), bo, *ln[2], o_in=o_in, ffn=(pk(W1), b1, pk(W2), b2, *ln[1], ff))
    assert rel_l2(got2, ref2) < 1e-2, rel_l2(got2, ref2)
    # no query positions (FACT.trans): pre tables absent
    ref3, _ = _ref_token_layer(x, nhead, Wo, bo, ln[0], Win=Win, bin_=bin_, ffn=(W1, b1, W2, b2, *ln[1]))
    got3 = x.clone()
    ops.token_layer(got3, nhead, pk(Wo), bo, *ln[0], w_in=pk(Win), b_in=bin_, ffn=(pk(W1), b1, pk(W2), b2, *ln[1], ff))
    assert rel_l2(got3, ref3) < 1e-2, rel_l2(got3, ref3)
