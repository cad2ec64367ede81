"""Tiny driver for ncu: a few launches of the fused dilated residual layer at the bench shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402

BF = torch.bfloat16
B, T, F = int(os.environ.get('PB', 64)), 4096, 256
cg = int(os.environ.get('CG', 2))
dev = 'cuda'
xs = [torch.randn(B, T, F, device=dev).to(BF) for _ in range(2)]
w3 = (torch.randn(3, F, F, device=dev) * (3 * F) ** -0.5).to(BF)
w1 = (torch.randn(F, F, device=dev) * F ** -0.5).to(BF)
b3, b1 = torch.randn(F, device=dev), torch.randn(F, device=dev)
ln = torch.full((B,), T, dtype=torch.int32, device=dev)
for i in range(6):
    ops.tcn_layer(xs[i % 2], xs[1 - i % 2], w3, b3, w1, b1, 4, len=ln, cta_group=cg)
torch.cuda.synchronize()
print('ok')
