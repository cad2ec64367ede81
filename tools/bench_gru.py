"""Per-step latency of the bidirectional GRU recurrence: fp32 FFMA cluster kernel vs the tensor-core (mma.sync) kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402


def main():
    dev = 'cuda'
    Hh, slot = 256, 4096
    for B, S in [(16, 2048), (8, 2048), (64, 2048), (1, 2048)]:
        gi = torch.randn(B, slot, 6 * Hh, device=dev)
        w = [torch.randn(3 * Hh, Hh, device=dev) * Hh ** -0.5 for _ in range(2)]
        bb = [torch.randn(3 * Hh, device=dev) * 0.1 for _ in range(2)]
        out = torch.zeros(B, slot, 2 * Hh, device=dev, dtype=torch.bfloat16)
        ns = torch.full((B,), S, dtype=torch.int32, device=dev)
        for mma in (False, True):
            def run():
                ops.gru_bidir(gi, w[0], bb[0], w[1], bb[1], out, ns, relu=True, mma=mma)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f'B={B:3d} S={S} {"mma " if mma else "ffma"} {ms:8.3f} ms  {ms * 1e3 / S:6.3f} us/step', flush=True)


if __name__ == '__main__':
    main()
