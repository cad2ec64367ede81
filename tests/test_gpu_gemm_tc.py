"""tcgen05 / TMA GEMM (factk_gemm_tc) against a torch-CPU fp32 reference on bf16- (or tf32-) rounded operands."""
import pytest
import torch

from fact_clip_b200 import ops
from fact_clip_b200.ops import S

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BF = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize('B,slot,K,N,lens', [(1, 128, 64, 64, [128]), (2, 256, 256, 256, [256, 130]), (3, 384, 512, 512, [384, 1, 200]),
                                              (2, 128, 128, 75, [128, 77]), (2, 75, 256, 768, None)])
@pytest.mark.parametrize('ydt', [BF, torch.float32])
def test_plain_bf16(B, slot, K, N, lens, ydt):
    x = rnd(B, slot, K, seed=1).to(BF)
    w = rnd(N, K, seed=2, scale=K ** -0.5).to(BF)
    bias = rnd(N, seed=3)
    out = torch.full((B, slot, N + (8 - N % 8) % 8), 7.0, dtype=ydt, device=DEV)[:, :, :N]
    ln = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=DEV)
    ops.gemm([S(x.to(DEV), w.to(DEV))], N, out, len=ln, bias=bias.to(DEV), relu=True, tc=True)
    torch.cuda.synchronize()
    ref = torch.relu(x.float() @ w.float().t() + bias)
    for b in range(B):
        T = slot if lens is None else lens[b]
        assert rel(out[b, :T], ref[b, :T]) < (1e-2 if ydt == BF else 1e-5), (b, rel(out[b, :T], ref[b, :T]))
        assert bool((out[b, T:].float() == 7.0).all())


@pytest.mark.parametrize('d', [1, 8, 512])
def test_conv_taps_residual(d):
    B, slot, F, lens = 2, 1024, 256, [1024, 700]
    x = rnd(B, slot, F, seed=4).to(BF)
    for b, T in enumerate(lens):
        x[b, T:] = 0                      # contract: tail rows reachable by taps are zero
    w3 = rnd(3, F, F, seed=5, scale=(3 * F) ** -0.5).to(BF)
    b3 = rnd(F, seed=6)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    xd = x.to(DEV)
    h = torch.zeros(B, slot, F, dtype=BF, device=DEV)
    ops.gemm([S(xd, w3[k].to(DEV), off=(k - 1) * d) for k in range(3)], F, h, len=ln, bias=b3.to(DEV), relu=True, tc=True)
    w1 = rnd(F, F, seed=7, scale=F ** -0.5).to(BF)
    y = torch.zeros(B, slot, F, dtype=BF, device=DEV)
    ops.gemm([S(h, w1.to(DEV))], F, y, len=ln, res=xd, tc=True)
    torch.cuda.synchronize()
    for b, T in enumerate(lens):
        xf = x[b, :T].float()
        z = torch.zeros(d, F)
        xp = torch.cat([z, xf, z])
        hr = torch.relu(sum(xp[k * d:k * d + T] @ w3[k].float().t() for k in range(3)) + b3)
        assert rel(h[b, :T], hr) < 1e-2
        yr = h[b, :T].float().cpu() @ w1.float().t() + xf
        assert rel(y[b, :T], yr) < 1e-2
        assert bool((y[b, T:] == 0).all())


def test_tf32_input_projection():
    B, slot, K, N, lens = 2, 256, 2048, 256, [256, 100]
    x, w, bias = rnd(B, slot, K, seed=8), rnd(N, K, seed=9, scale=K ** -0.5), rnd(N, seed=10)
    out = torch.zeros(B, slot, N, dtype=BF, device=DEV)
    ops.gemm([S(x.to(DEV), w.to(DEV))], N, out, len=torch.tensor(lens, dtype=torch.int32, device=DEV), bias=bias.to(DEV), tc=True)
    torch.cuda.synchronize()
    ref = x @ w.t() + bias
    for b, T in enumerate(lens):
        assert rel(out[b, :T], ref[b, :T]) < 1e-2


def test_pre_addend_gather_and_pervideo_weights():
    B, slot, K, N = 2, 256, 128, 128
    x = rnd(B, slot, K, seed=11).to(BF)
    w = rnd(B, N, K, seed=12, scale=K ** -0.5).to(BF)       # per-video weights
    pre = rnd(B, 40, N, seed=13)
    idx = torch.randint(0, 40, (B, slot), generator=torch.Generator().manual_seed(14), dtype=torch.int32)
    out = torch.zeros(B, slot, N, device=DEV)
    ops.gemm([S(x.to(DEV), w.to(DEV))], N, out, relu=True, alpha=0.5, pre=pre.to(DEV), pre_idx=idx.to(DEV), tc=True)
    torch.cuda.synchronize()
    for b in range(B):
        ref = torch.relu(0.5 * (x[b].float() @ w[b].float().t()) + pre[b][idx[b].long()])
        assert rel(out[b], ref) < 1e-5


def test_many_tiles_persistent():
    """More tiles than SMs: exercises the stage ring wrap-around and both TMEM accumulators."""
    B, slot, K, N = 6, 4096, 256, 256
    x = rnd(B, slot, K, seed=15).to(BF)
    w = rnd(N, K, seed=16, scale=K ** -0.5).to(BF)
    out = torch.zeros(B, slot, N, dtype=BF, device=DEV)
    ops.gemm([S(x.to(DEV), w.to(DEV))], N, out, tc=True)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    assert rel(out, ref) < 1e-2


@pytest.mark.parametrize('B,slot,K,N,lens', [(2, 512, 256, 512, [512, 300]), (3, 384, 512, 256, [384, 1, 129]), (1, 128, 64, 128, [77]),
                                              (5, 4096, 512, 512, [4096, 4000, 3000, 4096, 77])])
@pytest.mark.parametrize('relu', [False, True])
def test_gemm_pair(B, slot, K, N, lens, relu):
    """CTA-pair bf16 GEMM (factk_gemm_pair): bias, ReLU, rows past the end untouched; strided input / output rows."""
    x = rnd(B, slot, K + 64, seed=41).to(BF)[:, :, :K]              # lda = K + 64
    w = rnd(N, K, seed=42, scale=K ** -0.5).to(BF)
    bias = rnd(N, seed=43)
    full = torch.full((B, slot, N + 128), 7.0, dtype=BF, device=DEV)
    out = full[:, :, :N]                                            # ldy = N + 128
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    xd = x.to(DEV)
    assert ops.gemm_pair_ok(xd, w.to(DEV), N, out)
    ops.gemm_pair(xd, w.to(DEV), N, out, len=ln, bias=bias.to(DEV), relu=relu)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + bias
    ref = torch.relu(ref) if relu else ref
    for b, T in enumerate(lens):
        assert rel(out[b, :T], ref[b, :T]) < 6e-3, (b, rel(out[b, :T], ref[b, :T]))
        assert bool((out[b, T:].float() == 7.0).all())
    assert bool((full[:, :, N:].float() == 7.0).all())


def test_gemm_pair_pre_gather():
    """sf_merge form: relu(frame W2^T + bias + (seg W1^T)[seg_label]) with the gathered term as a pre-activation addend."""
    B, slot, K, N, S = 2, 512, 512, 256, 40
    x = rnd(B, slot, K, seed=44).to(BF)
    w = rnd(N, K, seed=45, scale=K ** -0.5).to(BF)
    bias = rnd(N, seed=46)
    pre = rnd(B, slot, N, seed=47)
    idx = torch.randint(0, S, (B, slot), generator=torch.Generator().manual_seed(48), dtype=torch.int32)
    lens = [512, 200]
    out = torch.zeros(B, slot, N, dtype=BF, device=DEV)
    ops.gemm_pair(x.to(DEV), w.to(DEV), N, out, len=torch.tensor(lens, dtype=torch.int32, device=DEV), bias=bias.to(DEV), relu=True,
                  pre=pre.to(DEV), pre_idx=idx.to(DEV))
    torch.cuda.synchronize()
    for b, T in enumerate(lens):
        ref = torch.relu(x[b, :T].float() @ w.float().t() + bias + pre[b][idx[b, :T].long()])
        assert rel(out[b, :T], ref) < 6e-3
        assert bool((out[b, T:] == 0).all())
