"""Fixed vs variable cost of the token-side tensor-core GEMM launches (64 videos x 75 tokens)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402
from fact_clip_b200.ops import S  # noqa: E402

dev = 'cuda'
B, M = 64, 75


def t_one(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for K, N in [(32, 64), (64, 256), (256, 256), (512, 256), (256, 512), (256, 768), (512, 512)]:
    x = torch.randn(B, M, K, device=dev)
    y = torch.zeros(B, M, N, device=dev)
    w = torch.randn(N, K, device=dev)
    # chained: output feeds nothing, but consecutive launches on one stream serialise (each waits for the previous grid)
    us = t_one(lambda: ops.gemm([S(x, w)], N, y, tc=True))
    print(f'tf32 tcgen05  K={K:4d} N={N:4d}: {us:6.1f} us per launch (serialised in one stream)')
x = torch.randn(B, M, 256, device=dev)
y = torch.zeros(B, M, 256, device=dev)
w = torch.randn(256, device=dev)
b_ = torch.randn(256, device=dev)
print(f'layernorm 256: {t_one(lambda: ops.layernorm(x, w, b_, y)):6.1f} us')
q = torch.randn(B, M, 768, device=dev)
o = torch.zeros(B, M, 256, device=dev)
print(f'mha_tokens: {t_one(lambda: ops.mha_tokens(q[:, :, :256], q[:, :, 256:512], q[:, :, 512:], o, 8)):6.1f} us')
