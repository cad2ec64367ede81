"""Fused dilated residual layer (factk_tcn_layer) against a torch-CPU restatement of
DilatedResidualLayer.forward (reference models/basic.py:154-171) on the same bf16-rounded operands."""
import pytest
import torch

from fact_clip_b200 import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BF = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def reference(x, w3, b3, w1, b1, d, T):
    """x: [slot, F] bf16 (zero tail); returns the first T rows of the layer output in fp32 (hidden tile rounded to bf16
    like the kernel's shared-memory operand)."""
    F = x.shape[1]
    xf = x[:T].float()
    z = torch.zeros(d, F)
    xp = torch.cat([z, xf, z])
    h = torch.relu(sum(xp[k * d:k * d + T] @ w3[k].float().t() for k in range(3)) + b3).to(BF).float()
    return xf + h @ w1.float().t() + b1


def run_case(B, slot, F, lens, d, cg, seed=0):
    x = rnd(B, slot, F, seed=seed + 1).to(BF)
    for b, T in enumerate(lens):
        x[b, T:] = 0
    w3 = rnd(3, F, F, seed=seed + 2, scale=(3 * F) ** -0.5).to(BF)
    w1 = rnd(F, F, seed=seed + 3, scale=F ** -0.5).to(BF)
    b3, b1 = rnd(F, seed=seed + 4), rnd(F, seed=seed + 5)
    y = torch.full((B, slot, F), 7.0, dtype=BF, device=DEV)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    ops.tcn_layer(x.to(DEV), y, w3.to(DEV), b3.to(DEV), w1.to(DEV), b1.to(DEV), d, len=ln, cta_group=cg)
    torch.cuda.synchronize()
    for b, T in enumerate(lens):
        if T > 0:
            ref = reference(x[b], w3, b3, w1, b1, d, T)
            e = rel(y[b, :T], ref)
            assert e < 6e-3, (b, T, d, cg, e)
            assert float((y[b, :T].float().cpu() - ref).abs().max()) < 0.06 * float(ref.abs().max())
        assert bool((y[b, T:].float() == 7.0).all()), 'rows past the video end must not be written'
    return y


@pytest.mark.parametrize('cg', [1, 2])
@pytest.mark.parametrize('F', [256, 128])
@pytest.mark.parametrize('d', [1, 2, 16, 128, 512])
def test_fused_layer(cg, F, d):
    run_case(3, 1024, F, [1024, 700, 1], d, cg)


@pytest.mark.parametrize('cg', [1, 2])
def test_fused_layer_odd_tiles_and_short(cg):
    # slot = 3 x 128: the CTA pair's second tile is past the slot end for the last super tile (TMA zero fill)
    run_case(2, 384, 256, [384, 129], 4, cg)
    run_case(1, 128, 256, [5], 1, cg)
    run_case(2, 256, 128, [200, 256], 64, cg)


@pytest.mark.parametrize('cg', [1, 2])
def test_fused_layer_many_tiles_deterministic(cg):
    """More tiles than SMs (ring wrap-around, accumulator ping-pong); bit-identical re-runs."""
    a = run_case(5, 4096, 256, [4096, 4000, 3000, 4096, 77], 8, cg, seed=10)
    b = run_case(5, 4096, 256, [4096, 4000, 3000, 4096, 77], 8, cg, seed=10)
    assert torch.equal(a, b)


def test_fused_pair_matches_single_cta():
    """CTA-pair (cta_group::2) and single-CTA variants agree to bf16 rounding of the output."""
    a = run_case(2, 1024, 256, [1024, 900], 32, 1, seed=20)
    b = run_case(2, 1024, 256, [1024, 900], 32, 2, seed=20)
    assert rel(a[0], b[0]) < 2e-3
