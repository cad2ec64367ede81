"""Time the frame-level class-softmax splice (process_feature + TDU argmax) at the metric shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402

B, slot, H, C = 64, 4096, 512, 75
x = torch.randn(B, slot, H, device='cuda').to(torch.bfloat16)
cl = torch.empty(B, slot, C, device='cuda')
pred = torch.empty(B, slot, dtype=torch.int32, device='cuda')
ln = torch.full((B,), slot, dtype=torch.int32, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for i in range(13):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.softmax_splice(x, C, cl, pred, len=ln)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts = sorted(ts[3:])
us = ts[len(ts) // 2]
print(f'splice {B}x{slot} rows, C={C}: {us:.1f} us, {B * slot * (C * 8 + 4) / us / 1e3:.0f} GB/s algorithmic')
