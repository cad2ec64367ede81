"""Training step on the GPU: backward-pass kernels against torch restatements, and whole-model gradients against
(a) gradient fixtures computed by the UNMODIFIED reference (tests/golden/grad_*.pt, make_grad_golden.py) and
(b) autograd through the oracle at the shipped configurations' widths."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import fact_oracle as O  # noqa: E402
import loss_oracle as LO  # noqa: E402
from fact_clip_b200 import config as C  # noqa: E402
from fact_clip_b200 import ops  # noqa: E402
from fact_clip_b200.loss import MatchCriterion  # noqa: E402
from fact_clip_b200.models.blocks import FACT, FACT_CLIP  # noqa: E402
from fact_clip_b200.utils.synth import make_batch, make_text_embeddings  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------------ kernels
def _rows(B, slot, N, seed, lens=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, slot, N, generator=g)
    ln = torch.tensor(lens if lens is not None else [slot] * B, dtype=torch.int32)
    return x, ln


@pytest.mark.parametrize('B,slot,N,K,off,lens', [(2, 300, 70, 50, 0, [300, 131]), (1, 2500, 64, 96, -3, [2333]), (3, 128, 33, 17, 2, [1, 128, 77]),
                                                (2, 256, 128, 128, 0, None)])
def test_wgrad_and_colsum(B, slot, N, K, off, lens):
    dz, ln = _rows(B, slot, N, 1, lens)
    a, _ = _rows(B, slot, K, 2, lens)
    dw0 = torch.randn(N, K, generator=torch.Generator().manual_seed(3))
    ref, refb = dw0.double().clone(), torch.zeros(N, dtype=torch.double)
    per = torch.zeros(B, N, K, dtype=torch.double)
    for b in range(B):
        T = int(ln[b])
        sh = torch.zeros(T, K, dtype=torch.double)
        lo, hi = max(0, -off), min(T, T - off)
        if hi > lo:
            sh[lo:hi] = a[b, lo + off:hi + off].double()
        per[b] = dz[b, :T].double().t() @ sh
        ref += 0.5 * per[b]
        refb += dz[b, :T].double().sum(0)
    dzc, ac, lnc, dw = dz.to(DEV), a.to(DEV), ln.to(DEV), dw0.to(DEV)
    ops.wgrad(dzc, ac, N, K, dw, off=off, len=lnc, alpha=0.5, accumulate=True)
    assert rel(dw, ref) < 1e-5
    pv = torch.zeros(B, N, K, device=DEV)
    ops.wgrad(dzc, ac, N, K, pv, off=off, len=lnc, accumulate=False, per_video=True)
    assert rel(pv, per) < 1e-5
    again = torch.zeros(B, N, K, device=DEV)
    ops.wgrad(dzc, ac, N, K, again, off=off, len=lnc, accumulate=False, per_video=True)
    assert torch.equal(pv, again)                      # fixed summation order
    bs = torch.zeros(N, device=DEV)
    ops.colsum(dzc, N, bs, len=lnc)
    assert rel(bs, refb) < 1e-5


@pytest.mark.parametrize('B,slot,N,K,off,lens', [(1, 2048, 256, 256, 0, [2048]), (2, 1024, 256, 256, -4, [1000, 333]), (1, 16384, 256, 256, 512, [16000]),
                                                (2, 512, 512, 256, 1, [512, 65]), (1, 1280, 256, 2048, 0, [1200]), (3, 256, 64, 128, -1, [256, 1, 130]),
                                                (1, 384, 192, 64, 0, [300])])
def test_wgrad_tensor_core(B, slot, N, K, off, lens):
    """tcgen05 weight gradient (MN-major operands straight from the row-major activations) against fp64."""
    g = torch.Generator().manual_seed(5)
    dz = torch.zeros(B, slot, N)
    a = torch.zeros(B, slot, K)
    for b, T in enumerate(lens):
        dz[b, :T] = torch.randn(T, N, generator=g)
        a[b, :T] = torch.randn(T, K, generator=g)
    dz16, a16 = dz.to(torch.bfloat16), a.to(torch.bfloat16)
    ref = torch.zeros(N, K, dtype=torch.double)
    per = torch.zeros(B, N, K, dtype=torch.double)
    for b, T in enumerate(lens):
        sh = torch.zeros(T, K, dtype=torch.double)
        lo, hi = max(0, -off), min(T, T - off)
        if hi > lo:
            sh[lo:hi] = a16[b, lo + off:hi + off].double()
        per[b] = dz16[b, :T].double().t() @ sh
        ref += per[b]
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    dw = torch.zeros(N, K, device=DEV)
    n0 = ops.COUNTERS['launches']
    ops.wgrad(dz16.to(DEV), a16.to(DEV), N, K, dw, off=off, len=ln, accumulate=False, tc=True)
    assert rel(dw, ref) < 1e-5, rel(dw, ref)
    pv = torch.zeros(B, N, K, device=DEV)
    ops.wgrad(dz16.to(DEV), a16.to(DEV), N, K, pv, off=off, len=ln, accumulate=False, per_video=True, tc=True)
    assert rel(pv, per) < 1e-5
    dw2 = dw.clone()
    ops.wgrad(dz16.to(DEV), a16.to(DEV), N, K, dw2, off=off, len=ln, alpha=0.5, accumulate=True, tc=True)
    assert rel(dw2, 1.5 * ref) < 1e-5


@pytest.mark.parametrize('B,T,M,nhead,dh,lens,dt', [(2, 700, 75, 8, 32, [700, 333], torch.float32), (1, 5000, 300, 8, 32, [4801], torch.bfloat16),
                                                    (3, 130, 9, 2, 5, [130, 1, 64], torch.float32), (1, 300, 300, 4, 64, [300], torch.float32)])
def test_heads_mm_three_layouts(B, T, M, nhead, dh, lens, dt):
    """ops.heads_mm against einsum in float64: the logits / apply / transposed-apply layouts of the attention cores
    (train.py mha_self, cross_attn), ragged videos, unaligned sizes, bf16 row operands, accumulate and the k-split path."""
    g = torch.Generator().manual_seed(B * 1000 + T)
    A, Mp = nhead * dh, (M + 3) // 4 * 4
    kk = torch.randn(B, T, A, generator=g).to(dt).to(DEV)
    q = torch.randn(B, M, A, generator=g).to(DEV)
    P = torch.randn(B, T, nhead * Mp, generator=g).to(DEV)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    valid = (torch.arange(T, device=DEV)[None] < ln[:, None]).double()[..., None, None]
    k64 = kk.double().view(B, T, nhead, dh) * valid
    q64 = q.double().view(B, M, nhead, dh)
    P64 = P.double().view(B, T, nhead, Mp)[..., :M] * valid
    # logits: L[b, t, h, m] = alpha sum_d kk q  (rows beyond len untouched)
    L_ = torch.full((B, T, nhead * Mp), 7.0, device=DEV)
    ops.heads_mm(kk, q, L_, T, M, dh, nhead, dh, dh, Mp, len=ln, len_mode=1, alpha=0.5)
    ref = 0.5 * torch.einsum('bthd,bmhd->bthm', k64, q64)
    got = L_.view(B, T, nhead, Mp).double()
    assert rel(got[..., :M] * valid, ref) < 1e-5
    assert torch.equal(got[..., M:], torch.full_like(got[..., M:], 7.0))
    for b in range(B):
        assert torch.equal(L_[b, lens[b]:], torch.full_like(L_[b, lens[b]:], 7.0))
    # apply along the tokens: dkk[b, t, h, d] += sum_m P q
    out = torch.ones(B, T, A, device=DEV, dtype=dt)
    ops.heads_mm(P, q, out, T, dh, M, nhead, Mp, dh, dh, b_kmajor=True, len=ln, len_mode=1, accumulate=True)
    ref = torch.einsum('bthm,bmhd->bthd', P64, q64) + valid
    assert rel(out.view(B, T, nhead, dh).double() * valid, ref) < (1e-5 if dt == torch.float32 else 6e-3)
    # reduction over the frames (k split when T is large): o[b, m, h, d] = sum_t P kk, then accumulated once more
    o = torch.zeros(B, M, A, device=DEV)
    for acc in (False, True):
        ops.heads_mm(P, kk, o, M, dh, T, nhead, Mp, dh, dh, a_kmajor=True, b_kmajor=True, len=ln, len_mode=2, alpha=0.25, accumulate=acc)
    ref = 0.5 * torch.einsum('bthm,bthd->bmhd', P64, k64)
    assert rel(o.view(B, M, nhead, dh), ref) < 1e-5
    o2 = torch.zeros_like(o)
    for acc in (False, True):
        ops.heads_mm(P, kk, o2, M, dh, T, nhead, Mp, dh, dh, a_kmajor=True, b_kmajor=True, len=ln, len_mode=2, alpha=0.25, accumulate=acc)
    assert torch.equal(o, o2)                    # fixed-order reduction: bit-reproducible


def test_row_kernels_backward():
    B, slot, H, Cc = 2, 200, 48, 7
    x, ln = _rows(B, slot, H, 5, [200, 99])
    xg = x.clone().requires_grad_(True)
    w, b = torch.randn(H, generator=torch.Generator().manual_seed(6)), torch.randn(H, generator=torch.Generator().manual_seed(7))
    wg, bg = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    r, _ = _rows(B, slot, H, 8, [200, 99])
    gy, _ = _rows(B, slot, H, 9, [200, 99])
    mask = (torch.arange(slot)[None, :] < ln[:, None]).float()[..., None]
    # layer norm (+ residual, + relu)
    y = torch.relu(torch.nn.functional.layer_norm(xg + r, (H,), wg, bg))
    (y * gy * mask).sum().backward()
    dv, dw, db = torch.zeros(B, slot, H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.layernorm_bwd(x.to(DEV), w.to(DEV), b.to(DEV), gy.to(DEV), dv, dw, db, res=r.to(DEV), relu=True, len=ln.to(DEV))
    assert rel(dv * mask.to(DEV), xg.grad * mask) < 1e-5 and rel(dw, wg.grad) < 1e-5 and rel(db, bg.grad) < 1e-5
    # softmax splice
    xg = x.clone().requires_grad_(True)
    feat = torch.cat([xg[..., :H - Cc], torch.softmax(xg[..., H - Cc:], -1)], -1)
    gcl = torch.randn(B, slot, Cc, generator=torch.Generator().manual_seed(10))
    ((feat * gy).sum(-1, keepdim=True) * mask + (xg[..., H - Cc:] * gcl).sum(-1, keepdim=True) * mask).sum().backward()
    dx = torch.zeros(B, slot, H, device=DEV)
    ops.splice_bwd(feat.detach().to(DEV), gy.to(DEV), gcl.to(DEV), dx, H, Cc, len=ln.to(DEV))
    assert rel(dx * mask.to(DEV), xg.grad) < 1e-5
    # l2 normalise
    xg = x.clone().requires_grad_(True)
    (torch.nn.functional.normalize(xg, dim=-1) * gy * mask).sum().backward()
    dx = torch.zeros(B, slot, H, device=DEV)
    ops.l2norm_bwd(x.to(DEV), gy.to(DEV), dx, len=ln.to(DEV))
    assert rel(dx * mask.to(DEV), xg.grad) < 1e-5
    # row softmax / column softmax
    M = 13
    lg, _ = _rows(B, slot, 16, 11, [200, 99])
    gp, _ = _rows(B, slot, 16, 12, [200, 99])
    lgg = lg.clone().requires_grad_(True)
    (torch.softmax(lgg[..., :M], -1) * gp[..., :M] * mask).sum().backward()
    P = torch.softmax(lg[..., :M], -1)
    Pp = torch.zeros(B, slot, 16)
    Pp[..., :M] = P
    dl = torch.zeros(B, slot, 16, device=DEV)
    ops.row_softmax_bwd(Pp.to(DEV), gp.to(DEV), dl, M, len=ln.to(DEV))
    assert rel(dl[..., :M] * mask.to(DEV), lgg.grad[..., :M]) < 1e-5
    lgg = lg.clone().requires_grad_(True)
    tot = 0
    for b_ in range(B):
        T = int(ln[b_])
        tot = tot + (torch.softmax(0.7 * lgg[b_, :T, :M], 0) * gp[b_, :T, :M]).sum()
    tot.backward()
    Pc = torch.zeros(B, slot, 16, device=DEV)
    ops.col_softmax(lg.to(DEV), Pc, M, scale=0.7, len=ln.to(DEV))
    for b_ in range(B):
        T = int(ln[b_])
        assert rel(Pc[b_, :T, :M], torch.softmax(0.7 * lg[b_, :T, :M], 0)) < 1e-5
    dl = torch.zeros(B, slot, 16, device=DEV)
    ops.col_softmax_bwd(Pc, gp.to(DEV), dl, M, scale=0.7, len=ln.to(DEV))
    assert rel(dl[..., :M] * mask.to(DEV), lgg.grad[..., :M]) < 1e-5


@pytest.mark.parametrize('dt,M,slot,lens', [(torch.bfloat16, 2432, 640, [640, 333]), (torch.float32, 300, 1500, [1500]), (torch.bfloat16, 75, 256, [9, 256])])
def test_col_softmax_dtypes_and_vector_path(dt, M, slot, lens):
    """Column softmax (softmax over the frames) forward / backward with bf16 tensors (the tcgen05 cross-attention path keeps its
    probability tensors in bf16) and the 4-columns-per-thread vector path, against float64 torch."""
    B = len(lens)
    g = torch.Generator().manual_seed(M + slot)
    lg = (torch.randn(B, slot, M, generator=g) * 2).to(dt)
    gp = torch.randn(B, slot, M, generator=g).to(dt)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    P = torch.zeros(B, slot, M, dtype=dt, device=DEV)
    ops.col_softmax(lg.to(DEV), P, M, scale=0.5, len=ln)
    dl = torch.zeros(B, slot, M, dtype=dt, device=DEV)
    ops.col_softmax_bwd(P, gp.to(DEV), dl, M, scale=0.5, len=ln)
    tol = 1e-5 if dt == torch.float32 else 6e-3
    for b_, T in enumerate(lens):
        x = lg[b_, :T].double().requires_grad_(True)
        ref = torch.softmax(0.5 * x, 0)
        assert rel(P[b_, :T], ref.detach()) < tol
        (P[b_, :T].double().cpu() * gp[b_, :T].double()).sum()             # the kernel differentiates through ITS rounded P
        (ref * gp[b_, :T].double()).sum().backward()
        assert rel(dl[b_, :T], x.grad) < 4 * tol
        assert float(P[b_, T:].float().abs().sum()) == 0.0 and float(dl[b_, T:].float().abs().sum()) == 0.0


def test_col_lse_chunked_rows():
    """factk_col_lse on a long rows tensor (512-row chunks across CTAs + fixed-order combine) equals torch.logsumexp, with the row mask."""
    B, slot, M = 2, 5000, 300
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(B, slot, M, generator=g) * 3).to(DEV)
    nrows = torch.tensor([5000, 1777], dtype=torch.int32, device=DEV)
    out = torch.zeros(B, M, device=DEV)
    ops.col_lse(x, M, nrows, out)
    for b_ in range(B):
        assert rel(out[b_], torch.logsumexp(x[b_, :int(nrows[b_])].double(), 0)) < 1e-6
    lab = torch.randint(0, 7, (B, slot), generator=g).to(torch.int32).to(DEV)
    cmap = torch.tensor([0, -1, 1, 2, -1, 3, 4], dtype=torch.int32, device=DEV)
    out2 = torch.zeros(B, M, device=DEV)
    ops.col_lse(x, M, nrows, out2, rmask0=lab, rmap=cmap)
    for b_ in range(B):
        keep = (cmap[lab[b_, :int(nrows[b_])].long()] >= 0)
        assert rel(out2[b_], torch.logsumexp(x[b_, :int(nrows[b_])][keep].double(), 0)) < 1e-6
    again = torch.zeros(B, M, device=DEV)
    ops.col_lse(x, M, nrows, again, rmask0=lab, rmap=cmap)
    assert torch.equal(out2, again)


@pytest.mark.parametrize('Hh,lens', [(32, [40, 7, 1]), (256, [300, 120])])
def test_gru_backward(Hh, lens):
    """BPTT kernel + the generic kernels around it against autograd through torch's nn.GRU."""
    B, slot, H = len(lens), 384, 2 * Hh
    torch.manual_seed(3)
    gru = torch.nn.GRU(H, Hh, 1, bidirectional=True)
    xs = [torch.randn(T, 1, H) for T in lens]
    gys = [torch.randn(T, 1, H) for T in lens]
    for p in gru.parameters():
        p.grad = None
    dxs = []
    for x, gy in zip(xs, gys):
        xx = x.clone().requires_grad_(True)
        y, _ = gru(xx)
        (y * gy).sum().backward()
        dxs.append(xx.grad[:, 0])
    sd = {k: v.detach().to(DEV) for k, v in gru.state_dict().items()}
    x = torch.zeros(B, slot, H, device=DEV)
    dy = torch.zeros(B, slot, H, device=DEV)
    for b, (a, g) in enumerate(zip(xs, gys)):
        x[b, :lens[b]] = a[:, 0].to(DEV)
        dy[b, :lens[b]] = g[:, 0].to(DEV)
    ns = torch.tensor(lens, dtype=torch.int32, device=DEV)
    Wih = torch.cat([sd['weight_ih_l0'], sd['weight_ih_l0_reverse']], 0)
    bih = torch.cat([sd['bias_ih_l0'], sd['bias_ih_l0_reverse']], 0)
    gi = torch.zeros(B, slot, 6 * Hh, device=DEV)
    ops.gemm([ops.S(x, Wih)], 6 * Hh, gi, len=ns, bias=bih)
    y = torch.zeros(B, slot, H, device=DEV)
    ops.gru_bidir(gi, sd['weight_hh_l0'], sd['bias_hh_l0'], sd['weight_hh_l0_reverse'], sd['bias_hh_l0_reverse'], y, ns, relu=False)
    gh = torch.zeros(B, slot, 6 * Hh, device=DEV)
    ops.gemm([ops.S(y[:, :, :Hh], sd['weight_hh_l0'], off=-1)], 3 * Hh, gh[:, :, :3 * Hh], len=ns, bias=sd['bias_hh_l0'])
    ops.gemm([ops.S(y[:, :, Hh:], sd['weight_hh_l0_reverse'], off=1)], 3 * Hh, gh[:, :, 3 * Hh:], len=ns, bias=sd['bias_hh_l0_reverse'])
    dgi, dgh = torch.zeros_like(gi), torch.zeros_like(gi)
    ops.gru_bwd(gi, gh, y, dy, sd['weight_hh_l0'], sd['weight_hh_l0_reverse'], dgi, dgh, ns)
    dx = torch.zeros(B, slot, H, device=DEV)
    ops.gemm([ops.S(dgi, Wih.t().contiguous())], H, dx, len=ns)
    for b in range(B):
        assert rel(dx[b, :lens[b]], dxs[b]) < 2e-5, b
    dWf, dWb = torch.zeros(3 * Hh, Hh, device=DEV), torch.zeros(3 * Hh, Hh, device=DEV)
    ops.wgrad(dgh[:, :, :3 * Hh], y[:, :, :Hh], 3 * Hh, Hh, dWf, off=-1, len=ns)
    ops.wgrad(dgh[:, :, 3 * Hh:], y[:, :, Hh:], 3 * Hh, Hh, dWb, off=1, len=ns)
    assert rel(dWf, gru.weight_hh_l0.grad) < 2e-5 and rel(dWb, gru.weight_hh_l0_reverse.grad) < 2e-5
    dbf = torch.zeros(3 * Hh, device=DEV)
    ops.colsum(dgh[:, :, :3 * Hh], 3 * Hh, dbf, len=ns)
    assert rel(dbf, gru.bias_hh_l0.grad) < 2e-5
    dWi = torch.zeros(6 * Hh, H, device=DEV)
    ops.wgrad(dgi, x, 6 * Hh, H, dWi, len=ns)
    assert rel(dWi, torch.cat([gru.weight_ih_l0.grad, gru.weight_ih_l0_reverse.grad], 0)) < 2e-5


def test_dropout_mask_is_reproducible_and_unbiased():
    B, slot, N, p = 2, 512, 256, 0.3
    x = torch.ones(B, slot, N, device=DEV)
    y1, y2, y3 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    ops.ew(ops.EW_DROPOUT, x, y1, N, p=p, seed=11, site=4)
    ops.ew(ops.EW_DROPOUT, x, y2, N, p=p, seed=11, site=4)
    ops.ew(ops.EW_DROPOUT, x, y3, N, p=p, seed=11, site=5)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)
    keep = (y1 > 0).float().mean().item()
    assert abs(keep - (1 - p)) < 5e-3 and abs(y1.mean().item() - 1.0) < 1e-2
    ch = torch.empty_like(x)
    ops.ew(ops.EW_DROPOUT_CH, x, ch, N, p=p, seed=11, site=6)
    assert torch.equal(ch[:, 0], ch[:, 100])            # one decision per (video, channel): every frame sees the same mask
    assert not torch.equal(ch[0, 0], ch[1, 0])


# ------------------------------------------------------------------------------------------------ whole model
def grad_names():
    return sorted(f[:-3] for f in os.listdir(GOLDEN) if f.startswith('grad_') and f.endswith('.pt'))


def _criterion(cfg, C_, bg):
    return MatchCriterion(cfg, C_, bg)


def _check_grads(net, ref_grads, tol, floor):
    worst, bad = 0.0, []
    gref = torch.sqrt(sum((g.double() ** 2).sum() for g in ref_grads.values()))
    for n, p in net.named_parameters():
        assert n in ref_grads, f'{n}: the reference has no gradient for it'
        assert p.grad is not None, f'{n}: no gradient (the reference produces one)'
        g, r = p.grad.double().cpu(), ref_grads[n].double()
        err = float((g - r).norm())
        # relative to the parameter's own gradient norm; parameters whose gradient is tiny against the global norm are
        # bounded in absolute terms instead
        denom = max(float(r.norm()), floor * float(gref))
        worst = max(worst, err / denom)
        if err / denom >= tol:
            bad.append((n, err / denom, float(r.norm())))
    assert not bad, f'{len(bad)} parameter gradients off: {bad[:8]}'
    return worst


@pytest.mark.parametrize('name', grad_names())
def test_gradients_vs_reference_fixtures(name):
    """fp32 mode: the gradient of every parameter within 1e-3 relative L2 of what the unmodified reference's
    loss.backward() produced (and the loss value itself)."""
    gr = torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)
    g = load_golden(gr['fixture'])
    cfg = C.tiny(**g['tiny_kwargs'])
    for k, v in gr['loss'].items():
        cfg.Loss[k] = v
    cfg.holdout_classes = list(gr['holdout'])
    cfg.CLIP.projection_dropout = 0.0
    net = (FACT_CLIP(cfg, g['in_dim'], g['n_classes'], make_text_embeddings(g['n_classes'])) if g['clip']
           else FACT(cfg, g['in_dim'], g['n_classes']))
    net.load_state_dict(g['state_dict'], strict=False)
    net.compute_mode = 'fp32'
    net = net.to(DEV).train()
    net.mcriterion = _criterion(cfg, g['n_classes'], gr['bg_ids'])
    vids = [g['videos'][i] for i in gr['videos']]
    loss, saves = net([v['x'].to(DEV) for v in vids], [v['label'].to(DEV) for v in vids], compute_loss=True)
    assert abs(float(loss.detach()) - gr["batch_loss"]) <= 1e-4 * abs(gr["batch_loss"])
    loss.backward()
    worst = _check_grads(net, gr['grads'], 1e-3, 1e-4)
    print(name, 'worst relative gradient error', worst)
    # a second step gives the same gradients bit for bit (fixed summation order, no atomics)
    g1 = {n: p.grad.clone() for n, p in net.named_parameters()}
    net.zero_grad()
    loss2, _ = net([v['x'].to(DEV) for v in vids], [v['label'].to(DEV) for v in vids], compute_loss=True)
    loss2.backward()
    assert all(torch.equal(g1[n], p.grad) for n, p in net.named_parameters())


def _oracle_grads(net_cpu, cfg, ncls, xs, ys, clip, bg, holdout):
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k in dict(net_cpu.named_parameters()))
          for k, v in net_cpu.state_dict().items()}
    hp = O.hparams_from_cfg(cfg, xs[0].shape[1], ncls)
    lp = LO.loss_params(cfg, bg)
    lp['holdout'] = list(holdout)
    total = 0
    for x, y in zip(xs, ys):
        o = O.forward_video(sd, hp, x, clip=clip)
        total = total + LO.loss_video(o, hp, y, lp, text_embeddings=sd.get('text_embeddings') if clip else None)['loss']
    total = total / len(xs)
    total.backward()
    return float(total), {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}


@pytest.mark.parametrize('preset,ncls,lens', [('havid_view0_lh_pt_holdout', 75, [300, 170]), ('epic_shape', 98, [400]), ('gtea', 11, [260])])
def test_gradients_vs_oracle_autograd_shipped_widths(preset, ncls, lens):
    """The shipped configurations' widths (F=A=256, M=75 / 300, MSTCN and MSTCN++, fpos, CLIP head, o2m matching) on short
    videos: loss and every parameter gradient against torch autograd through the oracle (fp32 mode, dropouts off)."""
    cfg = C.PRESETS[preset]()
    for blk in (cfg.Bi, cfg.Bu, cfg.BU):
        if blk.dropout is not None:
            blk.dropout = 0.0
    cfg.FACT.cmr, cfg.TM.use, cfg.CLIP.projection_dropout = 0.0, False, 0.0
    clip = bool(cfg.use_clip)
    holdout = list(cfg.holdout_classes) if clip else []
    torch.manual_seed(0)
    net = (FACT_CLIP(cfg, 2048, ncls, make_text_embeddings(ncls)) if clip else FACT(cfg, 2048, ncls))
    xs, ys = make_batch(lens, 2048, ncls, base_seed=60, nseg=6)
    ref_loss, ref = _oracle_grads(net, cfg, ncls, xs, ys, clip, [0], holdout)
    net.compute_mode = 'fp32'
    net = net.to(DEV).train()
    net.mcriterion = _criterion(cfg, ncls, [0])
    loss, _ = net([x.to(DEV) for x in xs], [y.to(DEV) for y in ys], compute_loss=True)
    assert abs(float(loss) - ref_loss) <= 2e-4 * abs(ref_loss), (float(loss), ref_loss)
    loss.backward()
    worst = _check_grads(net, ref, 2e-3, 1e-4)
    print(preset, 'loss', float(loss), 'worst relative gradient error', worst)


@pytest.mark.parametrize('preset,ncls,lens', [('havid_view0_lh_pt_holdout', 75, [384, 200]), ('epic_shape', 98, [512])])
def test_bf16_training_step_vs_oracle(preset, ncls, lens):
    """bf16 mode (tcgen05 forward / data-gradient GEMMs, tcgen05 weight gradients, bf16 activations and activation gradients,
    fp32 master gradients) with the oracle's segmentation teacher-forced: loss within 2e-2, the whole gradient within 2e-2
    relative L2 of fp32 autograd through the oracle, no parameter off by more than 10 % of its own gradient norm."""
    cfg = C.PRESETS[preset]()
    for blk in (cfg.Bi, cfg.Bu, cfg.BU):
        if blk.dropout is not None:
            blk.dropout = 0.0
    cfg.FACT.cmr, cfg.TM.use, cfg.CLIP.projection_dropout = 0.0, False, 0.0
    clip = bool(cfg.use_clip)
    holdout = list(cfg.holdout_classes) if clip else []
    torch.manual_seed(0)
    net = (FACT_CLIP(cfg, 2048, ncls, make_text_embeddings(ncls)) if clip else FACT(cfg, 2048, ncls))
    xs, ys = make_batch(lens, 2048, ncls, base_seed=60, nseg=6)
    ref_loss, ref = _oracle_grads(net, cfg, ncls, xs, ys, clip, [0], holdout)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    hp = O.hparams_from_cfg(cfg, 2048, ncls)
    with torch.no_grad():
        outs = [O.forward_video(sd, hp, x, clip=clip, fast_gru=True) for x in xs]
    nU = sum(1 for b in hp['blocks'] if b['type'] == 'U')
    forced = [[[b_['tdu_pred'] for b_ in o['blocks'] if 'tdu_pred' in b_][u].to(DEV) for o in outs] for u in range(nU)]
    net.compute_mode = 'bf16'
    net = net.to(DEV).train()
    net.mcriterion = _criterion(cfg, ncls, [0])
    loss, _ = net([x.to(DEV) for x in xs], [y.to(DEV) for y in ys], compute_loss=True, forced_preds=forced)
    assert abs(float(loss.detach()) - ref_loss) <= 2e-2 * abs(ref_loss), (float(loss.detach()), ref_loss)
    loss.backward()
    num = den = 0.0
    worst = ('', 0.0)
    gref = float(torch.sqrt(sum((g.double() ** 2).sum() for g in ref.values())))
    for n, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        g, r = p.grad.double().cpu(), ref[n].double()
        num += float(((g - r) ** 2).sum())
        den += float((r ** 2).sum())
        e = float((g - r).norm()) / max(float(r.norm()), 1e-2 * gref)
        if e > worst[1]:
            worst = (n, e)
    print(preset, 'loss', float(loss.detach()), ref_loss, 'global gradient rel-L2', (num / den) ** 0.5, 'worst parameter', worst)
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5
    assert worst[1] < 0.1, worst


def test_graph_captured_step_equals_eager():
    """net.train_graphs: forward and backward captured into CUDA graphs per batch shape; replayed steps give the same loss and
    bit-identical gradients as eager steps on new data of the same shape, including new dropout masks per step."""
    g = load_golden('tiny_m_iuU_clip_trained')
    cfg = C.tiny(**g['tiny_kwargs'])
    cfg.Loss.match, cfg.Loss.nullw, cfg.Loss.sw, cfg.Loss.pc = 'o2o', 0.1, 0.5, 0.2
    cfg.FACT.cmr = 0.25
    nets = []
    for graphs in (False, True):
        net = FACT_CLIP(cfg, g['in_dim'], g['n_classes'], make_text_embeddings(g['n_classes']))
        net.load_state_dict(g['state_dict'], strict=False)
        net.compute_mode = 'fp32'
        net = net.to(DEV).train()
        net.mcriterion = _criterion(cfg, g['n_classes'], [])
        net.train_graphs = graphs
        nets.append(net)
    vids = g['videos']
    batches = [[vids[0], vids[2]], [vids[2], vids[0]], [vids[0], vids[2]], [vids[2], vids[0]]]     # same (B, slot), other lengths order
    held = []           # losses of earlier steps stay referenced (as in a training loop that logs them): their graphs keep the
                        # parameters' AccumulateGrad nodes of the eager warm-up step alive -- the capture must not depend on those
    for step, batch in enumerate(batches):
        res = []
        for net in nets:
            net.zero_grad(set_to_none=True)
            loss, _ = net([v['x'].to(DEV) for v in batch], [v['label'].to(DEV) for v in batch], compute_loss=True)
            held.append(loss)
            loss.backward()
            res.append((float(loss.detach()), {n: p.grad.clone() for n, p in net.named_parameters()}))
        assert res[0][0] == res[1][0], (step, res[0][0], res[1][0])
        for n in res[0][1]:
            assert torch.equal(res[0][1][n], res[1][1][n]), (step, n)
    assert nets[1].train_engine()._cur.get('gB') is not None, 'the backward pass was never captured'


def test_train_mode_augmentations():
    """Channel masking (FACT.cmr), dropout and time masking change the train-mode forward, are regenerated identically in the
    backward pass (finite gradients for every parameter), and vanish in eval mode."""
    g = load_golden('tiny_m_iuU_clip')
    cfg = C.tiny(**g['tiny_kwargs'])
    cfg.FACT.cmr = 0.3
    for blk in (cfg.Bi, cfg.Bu, cfg.BU):
        blk.dropout = 0.2
    cfg.TM.use, cfg.TM.t, cfg.TM.m, cfg.TM.p = True, 10, 3, 0.2
    net = FACT_CLIP(cfg, g['in_dim'], g['n_classes'], make_text_embeddings(g['n_classes']))
    net.load_state_dict(g['state_dict'], strict=False)
    net.compute_mode = 'fp32'
    net = net.to(DEV)
    net.mcriterion = _criterion(cfg, g['n_classes'], [])
    xs = [v['x'].to(DEV) for v in g['videos']]
    ys = [v['label'].to(DEV) for v in g['videos']]
    net.eval()
    e1 = net(xs, ys, compute_loss=True)[0]
    net.train()
    import random
    random.seed(0)
    l1, _ = net(xs, ys, compute_loss=True)
    l1.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    random.seed(0)
    net.train_engine().step_no -= 1                      # same hash seed as the previous step
    l2, _ = net(xs, ys, compute_loss=True)
    assert float(l1) == float(l2) and abs(float(l1) - float(e1)) > 1e-4
    l3, _ = net(xs, ys, compute_loss=True)               # next step: new masks
    assert float(l3) != float(l1)
    net.eval()
    assert float(net(xs, ys, compute_loss=True)[0]) == float(e1)


def test_train_mode_input_augmentations_vs_oracle():
    """Row a5 (blocks.py:614-622): channel masking (nn.Dropout2d over whole feature channels, FACT.cmr) and time masking
    (basic.time_mask, spans from Python's ``random``).  The masks this step drew are read back -- the channel mask by running
    the same hashed dropout over a tensor of ones, the spans by replaying ``random`` from the same seed -- applied to the inputs
    on the host, and the oracle's loss and gradients on those masked inputs must equal the train-mode step's (network dropouts off)."""
    import random
    cfg = C.PRESETS['gtea']()
    for blk in (cfg.Bi, cfg.Bu, cfg.BU):
        if blk.dropout is not None:
            blk.dropout = 0.0
    cfg.FACT.cmr = 0.4
    cfg.TM.use, cfg.TM.t, cfg.TM.m, cfg.TM.p = True, 40, 4, 0.1
    ncls, D, lens = 11, 2048, [300, 170]
    torch.manual_seed(0)
    net = FACT(cfg, D, ncls)
    xs, ys = make_batch(lens, D, ncls, base_seed=61, nseg=6)
    cpu_net = net
    net.compute_mode = 'fp32'
    net = net.to(DEV).train()
    net.mcriterion = _criterion(cfg, ncls, [10])
    random.seed(1234)
    loss, _ = net([x.to(DEV) for x in xs], [y.to(DEV) for y in ys], compute_loss=True)
    loss.backward()
    eng = net.train_engine()
    # the channel mask of this step: site 1 (the first dropout of the forward), seed = engine seed + step counter
    ones = torch.ones(len(lens), 384, D, device=DEV)
    keep = torch.empty_like(ones)
    ops.ew(ops.EW_DROPOUT_CH, ones, keep, D, p=0.4, seed=eng.seed + eng.step_no, site=1)
    keep = keep[:, 0].cpu()                                  # [B, D]: 0 or 1 / (1 - p)
    assert 0.5 < float((keep > 0).float().mean()) < 0.7
    random.seed(1234)
    masked = []
    for b, (x, T) in enumerate(zip(xs, lens)):
        xm = x * keep[b][None, :]
        for t0, t1 in eng.time_mask_spans(T, cfg.TM):
            xm[t0:t1] = 0.0
        masked.append(xm)
    assert any(float((m.abs().sum(1) == 0).sum()) > 0 for m in masked), 'no frame was time-masked: the test would not see TM'
    ref_loss, ref = _oracle_grads(cpu_net.cpu(), cfg, ncls, masked, ys, False, [10], [])
    assert abs(float(loss.detach()) - ref_loss) <= 2e-4 * abs(ref_loss), (float(loss.detach()), ref_loss)
    net = net.to(DEV)
    worst = _check_grads(net, ref, 2e-3, 1e-4)
    print('train-mode augmentations vs oracle on the masked inputs: loss', float(loss.detach()), ref_loss, 'worst gradient error', worst)
