"""Where the end-to-end step time goes: duration of the H2D copies and of the graph replay while they overlap."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fact_clip_b200.utils.synth import make_video  # noqa: E402

dev = torch.device('cuda', 0)
net, cfg = bench.build_model('bf16')
net = net.to(dev)
eng = net.engine()
B, T, D = int(os.environ.get('PB', 64)), 4096, 2048
host = torch.empty(B, T, D, pin_memory=True)
v = make_video(T, D, 75, seed=1)[0]
for b in range(B):
    host[b].copy_(v)
seqs = [host[b] for b in range(B)]
for _ in range(3):
    net.submit(seqs, None).result()
torch.cuda.synchronize()
# copy only
cs = torch.cuda.Stream()
x = eng.buf('input_p0', (B, T, D), torch.float32)
with torch.cuda.stream(cs):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cs)
    for _ in range(3):
        for b in range(B):
            x[b].copy_(seqs[b], non_blocking=True)
    e1.record(cs)
torch.cuda.synchronize()
print(f'copy only (64 per-video copies): {e0.elapsed_time(e1) / 3:.2f} ms per batch')
# compute only
ln = torch.full((B,), T, dtype=torch.int32, device=dev)
e0.record()
for _ in range(3):
    eng.run_packed_graphed(x, ln, [T] * B)
e1.record()
torch.cuda.synchronize()
print(f'compute only (graph replay): {e0.elapsed_time(e1) / 3:.2f} ms per batch')
# both, explicitly overlapped, no host sync in between
e0.record()
with torch.cuda.stream(cs):
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(cs)
    for _ in range(3):
        for b in range(B):
            x[b].copy_(seqs[b], non_blocking=True)
    c1.record(cs)
x2 = eng.buf('input', (B, T, D), torch.float32)
k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.run_packed_graphed(x2, ln, [T] * B)
k0.record()
for _ in range(6):
    eng.run_packed_graphed(x2, ln, [T] * B)
k1.record()
torch.cuda.synchronize()
print(f'overlapped: copies {c0.elapsed_time(c1) / 3:.2f} ms per batch, compute {k0.elapsed_time(k1) / 6:.2f} ms per batch')
# the pipelined API
t0 = time.perf_counter()
pend = []
n = 6
for i in range(n):
    pend.append(net.submit(seqs, None))
    if len(pend) > 1:
        pend.pop(0).result()
while pend:
    pend.pop(0).result()
dt = (time.perf_counter() - t0) / n
print(f'submit/result pipeline: {dt * 1e3:.2f} ms per batch')
# host-side cost of submit alone
torch.cuda.synchronize()
t0 = time.perf_counter()
h = net.submit(seqs, None)
t1 = time.perf_counter()
h.result()
print(f'host time inside submit(): {(t1 - t0) * 1e3:.2f} ms')
