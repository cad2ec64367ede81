"""segment_mean kernel time for short-segment (S ~ 0.6 T) and long-segment (S small) segmentations, 64 x 4096 x 512 bf16."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402

dev = 'cuda'
B, T, E = 64, 4096, 512
x = torch.randn(B, T, E, device=dev).to(torch.bfloat16)
seg = torch.zeros(B, T, E, device=dev, dtype=torch.bfloat16)
I32 = torch.int32
bufs = [torch.zeros(B, T, dtype=I32, device=dev) for _ in range(4)]
nseg = torch.zeros(B, dtype=I32, device=dev)
ln = torch.full((B,), T, dtype=I32, device=dev)
g = torch.Generator(device='cpu').manual_seed(0)
for name, mk in (('short runs (1..3)', lambda: torch.randint(1, 4, (T,), generator=g)),
                 ('mixed (1..400)', lambda: torch.randint(1, 401, (T,), generator=g)),
                 ('one segment', lambda: torch.full((T,), T))):
    pred = torch.zeros(B, T, dtype=I32)
    for b in range(B):
        runs = mk()
        lab = torch.repeat_interleave(torch.arange(T) % 7, runs)[:T]
        pred[b] = lab.to(I32)
    pred = pred.to(dev)
    ops.tdu_segment(pred, bufs[0], bufs[1], bufs[2], bufs[3], nseg, len=ln)
    ws = torch.empty(ops.segment_mean_ws(B, T, E), device=dev)
    for _ in range(3):
        ops.segment_mean(x, seg, bufs[0], bufs[1], bufs[2], nseg, ws=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.segment_mean(x, seg, bufs[0], bufs[1], bufs[2], nseg, ws=ws)
    e1.record()
    torch.cuda.synchronize()
    print(f'{name:20s} S in [{int(nseg.min())}, {int(nseg.max())}]: {e0.elapsed_time(e1) / 10 * 1e3:7.1f} us '
          f'(read {B * T * E * 2 / 1e6:.0f} MB -> {B * T * E * 2 / 1e6 / 6542.1 * 1e3:.0f} us at HBM peak)')
