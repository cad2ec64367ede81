"""Driver for ncu / timing of the two HBM-bound frame-side GEMMs at the bench shape (64 videos x 4096 frames):
  in-proj  : fp32 features [T, 2048] -> bf16 [T, 256]   (tf32 tcgen05, factk_gemm_tc)   algorithmic bytes 4*2048 + 2*256 per frame
  conv_out : bf16 [T, 256] -> bf16 [T, 512]             (CTA-pair tcgen05, factk_gemm_pair)  algorithmic bytes 2*256 + 2*512 per frame
Prints achieved GB/s against the algorithmic bytes (CUDA events, inputs far larger than L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402
from fact_clip_b200.ops import S  # noqa: E402

dev, BF = 'cuda', torch.bfloat16
B, T, D, F, H = 64, 4096, 2048, 256, 512
x = torch.randn(B, T, D, device=dev)
w_in = torch.randn(F, D, device=dev) * D ** -0.5
b_in = torch.randn(F, device=dev)
f = torch.zeros(B, T, F, dtype=BF, device=dev)
w_out = (torch.randn(H, F, device=dev) * F ** -0.5).to(BF)
b_out = torch.randn(H, device=dev)
y = torch.zeros(B, T, H, dtype=BF, device=dev)
ln = torch.full((B,), T, dtype=torch.int32, device=dev)


def in_proj():
    ops.gemm([S(x, w_in)], F, f, len=ln, bias=b_in, tc=True, tag='in_proj')


def conv_out():
    ops.gemm_pair(f, w_out, H, y, len=ln, bias=b_out, tag='conv_out')


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


for name, fn, bytes_per_frame in (('in_proj', in_proj, 4 * D + 2 * F), ('conv_out', conv_out, 2 * F + 2 * H)):
    s = timeit(fn)
    print(f'{name}: {s * 1e6:.1f} us, algorithmic {bytes_per_frame} B/frame x {B * T} frames = {bytes_per_frame * B * T / 1e6:.0f} MB '
          f'-> {bytes_per_frame * B * T / s / 1e9:.0f} GB/s')
