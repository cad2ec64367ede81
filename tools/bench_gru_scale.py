import os, sys, torch
sys.path.insert(0, '/root/repo')
from fact_clip_b200 import ops
dev='cuda'; Hh, slot = 256, 2304
for B in (8, 16, 24, 32, 40, 48, 56, 64):
    S = 2048
    gi = torch.randn(B, slot, 6 * Hh, device=dev)
    w = [torch.randn(3 * Hh, Hh, device=dev) * Hh ** -0.5 for _ in range(2)]
    bb = [torch.randn(3 * Hh, device=dev) * 0.1 for _ in range(2)]
    out = torch.zeros(B, slot, 2 * Hh, device=dev, dtype=torch.bfloat16)
    ns = torch.full((B,), S, dtype=torch.int32, device=dev)
    run = lambda: ops.gru_bidir(gi, w[0], bb[0], w[1], bb[1], out, ns, relu=True, mma=True)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f'B={B:3d} clusters={2*((B+7)//8):2d} {ms*1e3/S:6.3f} us/step', flush=True)
