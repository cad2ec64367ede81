"""Driver for ncu: the f2a column softmax + weighted row sum at the bench shape (64 x 4096 rows, 75 tokens, 512 channels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402

dev = 'cuda'
B, T, M, E = 64, 4096, 75, 512
logit = torch.randn(B, T, 76, device=dev) * 3
x = torch.randn(B, T, E, device=dev).to(torch.bfloat16)
out = torch.zeros(B, M, E, device=dev)
ws = torch.empty(ops.col_softmax_ws(B, T, M, E), device=dev)
ln = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.col_softmax_apply(logit, x, out, M, ws, len=ln)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.col_softmax_apply(logit, x, out, M, ws, len=ln)
e1.record()
torch.cuda.synchronize()
print(f'col_softmax_apply: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us')
