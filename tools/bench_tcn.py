"""Micro-benchmark of the fused dilated residual layer vs the two-GEMM path (CUDA events, L2 flushed by size)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402
from fact_clip_b200.ops import S  # noqa: E402

BF = torch.bfloat16


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = 'cuda'
    for B, T, F in [(16, 4096, 256), (64, 4096, 256), (16, 4096, 128)]:
        xs = [torch.randn(B, T, F, device=dev).to(BF) for _ in range(2)]
        w3 = (torch.randn(3, F, F, device=dev) * (3 * F) ** -0.5).to(BF)
        w1 = (torch.randn(F, F, device=dev) * F ** -0.5).to(BF)
        b3, b1 = torch.randn(F, device=dev), torch.randn(F, device=dev)
        ln = torch.full((B,), T, dtype=torch.int32, device=dev)
        tmp = torch.zeros(B, T, F, dtype=BF, device=dev)
        flops = 8.0 * F * F * B * T
        for d in (1, 16, 512):
            def unfused():
                ops.gemm([S(xs[0], w3[k], off=(k - 1) * d) for k in range(3)], F, tmp, len=ln, bias=b3, relu=True, tc=True)
                ops.gemm([S(tmp, w1)], F, xs[1], len=ln, bias=b1, res=xs[0], tc=True)
            us = timeit(unfused)
            print(f'B={B} T={T} F={F} d={d:3d} unfused   {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s', flush=True)
            for cg in (1, 2):
                us = timeit(lambda: ops.tcn_layer(xs[0], xs[1], w3, b3, w1, b1, d, len=ln, cta_group=cg))
                print(f'B={B} T={T} F={F} d={d:3d} fused cg={cg} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    main()
