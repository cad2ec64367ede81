"""Per-step timeline (clock64) of the Hh = 256 GRU BPTT kernel at batch 1.  Usage: python tools/prof_gru_bwd.py [S]"""
import sys

import torch

sys.path.insert(0, '.')
from fact_clip_b200 import _lib, ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev, Hh, B = 'cuda', 256, 1
g = torch.Generator().manual_seed(0)
mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
gi, gh = mk(B, S, 6 * Hh), mk(B, S, 6 * Hh)
y, dy = mk(B, S, 2 * Hh, sc=0.5), mk(B, S, 2 * Hh, sc=0.1)
wf, wb = mk(3 * Hh, Hh, sc=Hh ** -0.5), mk(3 * Hh, Hh, sc=Hh ** -0.5)
dgi, dgh = torch.zeros_like(gi), torch.zeros_like(gh)
nseg = torch.tensor([S], dtype=torch.int32, device=dev)
lib = _lib.load()
for _ in range(2):
    ops.gru_bwd(gi, gh, y, dy, wf, wb, dgi, dgh, nseg)
dbg = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
lib.factk_gru_bwd_debug(dbg.data_ptr())
ops.gru_bwd(gi, gh, y, dy, wf, wb, dgi, dgh, nseg)
torch.cuda.synchronize()
lib.factk_gru_bwd_debug(None)
t = dbg.view(64, 8).cpu()
names = ['gates', 'send', 'fetch', 'wait', 'matvec', 'barrier', 'carry -> next']
d = [(t[8:60, i + 1] - t[8:60, i]).float().mean().item() for i in range(6)] + [(t[9:61, 0] - t[8:60, 6]).float().mean().item()]
for n, v in zip(names, d):
    print(f'  {n:14s} {v:8.0f} cycles')
print(f'  step           {(t[9:61, 0] - t[8:60, 0]).float().mean().item():8.0f} cycles')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.gru_bwd(gi, gh, y, dy, wf, wb, dgi, dgh, nseg)
e1.record()
torch.cuda.synchronize()
print(f'  {e0.elapsed_time(e1) / 5 * 1e3 / S:.3f} us per step (S = {S})')
