"""Does grouping videos of similar segment count into the same 8-video chain group shorten the GRU launch?  64 videos with
segment counts spread like the bench workload (564..2572): (a) in random order (every group holds a long video, all 16
clusters stay active to the end), (b) sorted by length (groups finish one after the other), (c) all equal to the maximum."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402


def main():
    dev = 'cuda'
    Hh, slot, B = 256, 4096, 64
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(564, 2573, (B,), generator=g)
    lens[0] = 2572
    cases = {'random order': lens, 'sorted': torch.sort(lens, descending=True).values, 'all 2572': torch.full((B,), 2572)}
    gi = torch.randn(B, slot, 6 * Hh, device=dev)
    w = [torch.randn(3 * Hh, Hh, device=dev) * Hh ** -0.5 for _ in range(2)]
    bb = [torch.randn(3 * Hh, device=dev) * 0.1 for _ in range(2)]
    out = torch.zeros(B, slot, 2 * Hh, device=dev, dtype=torch.bfloat16)
    for name, l in cases.items():
        ns = l.to(torch.int32).to(dev)
        run = lambda: ops.gru_bidir(gi, w[0], bb[0], w[1], bb[1], out, ns, relu=True, mma=True)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'{name:14s} {ms:7.3f} ms  ({ms * 1e3 / 2572:.3f} us per step of the longest chain)', flush=True)


if __name__ == '__main__':
    main()
