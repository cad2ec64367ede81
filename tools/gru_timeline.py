"""clock64 timeline of one CTA of the tensor-core GRU kernel (per step: wait, mma, sync, gates+send, tail)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import _lib as L  # noqa: E402

dev = 'cuda'
Hh, slot = 256, 4096
B = int(os.environ.get('PB', 16))
S = 512
gi = torch.randn(B, slot, 6 * Hh, device=dev)
w = [torch.randn(3 * Hh, Hh, device=dev) * Hh ** -0.5 for _ in range(2)]
bb = [torch.randn(3 * Hh, device=dev) * 0.1 for _ in range(2)]
out = torch.zeros(B, slot, 2 * Hh, device=dev, dtype=torch.bfloat16)
ns = torch.full((B,), S, dtype=torch.int32, device=dev)
dbg = torch.zeros(64, 16, dtype=torch.int64, device=dev)
for _ in range(2):
    L.call('factk_gru_bidir_mma_dbg', gi.data_ptr(), w[0].data_ptr(), bb[0].data_ptr(), w[1].data_ptr(), bb[1].data_ptr(), Hh,
           out.data_ptr(), 1, 2 * Hh, 1, B, slot, ns.data_ptr(), dbg.data_ptr(), L.stream())
torch.cuda.synchronize()
d = dbg.cpu()
print('step  begin  wait_done  mma+part  sync  gates+send  tail   (cycles, relative to step begin; begin = delta to previous step)')
for t in range(20, 32):
    for c in range(2):
        r = [int(x) for x in d[t, c * 8:c * 8 + 8]]
        if r[0] == 0:
            continue
        prev = int(d[t - 1, c * 8])
        print(f'{t:4d}{"AB"[c]} {r[0] - prev:6d} {r[1] - r[0]:9d} {r[2] - r[1]:9d} {r[3] - r[2]:5d} {r[4] - r[3]:10d} {r[5] - r[4]:6d}')
