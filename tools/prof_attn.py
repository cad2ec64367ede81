"""Driver for ncu / timing: the SCALayer cross-attention core at the bench shape (64 videos x 4096 frames, 75 tokens, 8 heads x 32)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402

dev = 'cuda'
B, T, M, nh, dh = int(os.environ.get('B', 64)), int(os.environ.get('T', 4096)), int(os.environ.get('M', 75)), 8, 32
E = nh * dh
q = torch.randn(B, M, E, device=dev)
kv = torch.randn(B, T, 2 * E, device=dev).to(torch.bfloat16)
o = torch.zeros(B, M, E, device=dev)
ws = torch.empty(ops.attn_rows_ws(B, T, M, nh, dh), device=dev)
ln = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.attn_rows(q, kv[..., :E], kv[..., E:], o, nh, ws, len=ln)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.attn_rows(q, kv[..., :E], kv[..., E:], o, nh, ws, len=ln)
e1.record()
torch.cuda.synchronize()
print(f'attn_rows (+combine) B={B} T={T} M={M}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us  '
      f'(K/V bytes {B * T * 2 * E * 2 / 1e6:.0f} MB, algorithmic FLOP {4 * B * M * T * E / 1e9:.1f} G)')

if os.environ.get('TIMELINE'):
    from fact_clip_b200 import _lib
    dbg = torch.zeros(512, dtype=torch.int64, device=dev)
    _lib.load().factk_attn_tc_debug(dbg.data_ptr())
    ops.attn_rows(q, kv[..., :E], kv[..., E:], o, nh, ws, len=ln)
    torch.cuda.synchronize()
    _lib.load().factk_attn_tc_debug(None)
    d = dbg.cpu().tolist()
    t0 = d[0]
    print('softmax group 0, warp 0 of CTA (0,0): cycles since loop start  [wait S begin, S ready, exp done, P buffer free]')
    prev = t0
    for n in range(32):
        a = [x - t0 for x in d[1 + 4 * n:5 + 4 * n]]
        print(f'  n={n:2d} head={(2 * n) % 8} tile={n // 4}: {a}  S wait {a[1] - a[0]:5d}  softmax {a[2] - a[1]:5d}  P wait {a[3] - a[2]:5d}  iteration {a[0] - (prev - t0):6d}')
        prev = d[1 + 4 * n]
    print('entry -> loop start', t0 - d[500], ' loop', d[501] - t0, ' wait o_full', d[502] - d[501], ' epilogue writes', d[503] - d[502])
    raise SystemExit
    print('MMA thread: [S: wait s_empty begin, issue] [PV: wait p_full begin, issue] (cycles since the softmax loop start)')
    for it in range(24):
        a = [x - t0 for x in d[300 + 4 * it:304 + 4 * it]]
        print(f'  it={it:2d} grp={it & 1}: S {a[0]:6d} -> {a[1]:6d}   PV {a[2]:6d} -> {a[3]:6d}')
