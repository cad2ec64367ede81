"""Pipeline timeline of the fused dilated residual layer (CTA 0): clock64 stamps per tile."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import _lib as L  # noqa: E402

BF = torch.bfloat16
B, T, F = int(os.environ.get('PB', 64)), 4096, 256
cg = int(os.environ.get('CG', 2))
dev = 'cuda'
xs = [torch.randn(B, T, F, device=dev).to(BF) for _ in range(2)]
w3 = (torch.randn(3, F, F, device=dev) * (3 * F) ** -0.5).to(BF)
w1 = (torch.randn(F, F, device=dev) * F ** -0.5).to(BF)
b3, b1 = torch.randn(F, device=dev), torch.randn(F, device=dev)
ln = torch.full((B,), T, dtype=torch.int32, device=dev)
dbg = torch.zeros(64, 16, dtype=torch.int64, device=dev)
for i in range(4):
    L.call('factk_tcn_layer_dbg', xs[0].data_ptr(), xs[1].data_ptr(), w3.data_ptr(), b3.data_ptr(), w1.data_ptr(), b1.data_ptr(),
           B, T, F, 4, ln.data_ptr(), cg, dbg.data_ptr(), L.stream())
torch.cuda.synchronize()
d = dbg.cpu()
t0 = int(d[0, 0])
names = ['g1_begin', 'g1_acc_free', 'g1_issued', 'g2_begin', 'g2_hready', 'g2_issued', 'e1_begin', 'e1_D1ready', 'e1_done',
         'e2_begin', 'e2_D2ready', 'e2_tmem_free', 'e2_done']
print('tile ' + ' '.join(f'{n:>12s}' for n in names))
for it in range(16):
    if int(d[it, 0]) == 0:
        break
    print(f'{it:4d} ' + ' '.join(f'{int(d[it, k]) - t0:12d}' for k in range(13)))
