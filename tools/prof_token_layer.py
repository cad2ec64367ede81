"""Stage timeline (clock64, CTA 0) and launch time of the fused token-layer kernel at the metric configuration's shape.
Usage: python tools/prof_token_layer.py [B M A nhead ff]"""
import sys

import torch

sys.path.insert(0, '.')
from fact_clip_b200 import _lib, ops  # noqa: E402

B, M, A, nh, ff = (int(a) for a in sys.argv[1:6]) if len(sys.argv) >= 6 else (64, 75, 256, 8, 512)
dev = 'cuda'
g = torch.Generator().manual_seed(0)
mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)
x, pos = mk(B, M, A), mk(M, A)
Win, bin_, Wo, bo = mk(3 * A, A, sc=A ** -0.5), mk(3 * A), mk(A, A, sc=A ** -0.5), mk(A)
W1, b1, W2, b2 = mk(ff, A, sc=A ** -0.5), mk(ff), mk(A, ff, sc=ff ** -0.5), mk(A)
lw, lb = 1 + mk(A, sc=0.1), mk(A, sc=0.1)
pk = ops.pack_token_weight
args = dict(w_in=pk(Win), b_in=bin_, pre_qk=(pos @ Win[:2 * A].t()).contiguous(), ffn=(pk(W1), b1, pk(W2), b2, lw, lb, ff))
wo = pk(Wo)
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
lib = _lib.load()
for _ in range(3):
    ops.token_layer(x.clone(), nh, wo, bo, lw, lb, **args)
lib.factk_token_layer_debug(dbg.data_ptr())
ops.token_layer(x.clone(), nh, wo, bo, lw, lb, **args)
torch.cuda.synchronize()
lib.factk_token_layer_debug(None)
t = dbg.cpu().tolist()
names = ['load rows', 'q|k|v GEMM', 'attention', 'out_proj', 'norm1', 'cq', 'linear1', 'linear2', 'norm2']
for i, n in enumerate(names):
    print(f'  {n:12s} {t[i + 1] - t[i]:8d} cycles')
print(f'  total        {t[9] - t[0]:8d} cycles')
xs = [x.clone() for _ in range(20)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for xx in xs:
    ops.token_layer(xx, nh, wo, bo, lw, lb, **args)
e1.record()
torch.cuda.synchronize()
print(f'  {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (B={B}, M={M}, A={A}, heads={nh}, ff={ff})')
