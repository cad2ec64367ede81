"""Host->device copy bandwidth from pinned memory (the ceiling of the end-to-end number: 8 KiB of fp32 features per frame)."""
import torch

dev = 'cuda'
for mib in (32, 512, 2048):
    n = mib * 2**20 // 4
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f'H2D {mib:5d} MiB pinned: {ms:8.3f} ms  {mib / 1024 / (ms * 1e-3):6.1f} GiB/s')
    h2 = torch.empty(n // 4, dtype=torch.float32, pin_memory=True)
    e0.record()
    for _ in range(5):
        h2.copy_(d[:n // 4], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f'D2H {mib // 4:5d} MiB pinned: {ms:8.3f} ms  {mib / 4 / 1024 / (ms * 1e-3):6.1f} GiB/s')
