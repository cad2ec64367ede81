"""The BASELINE.json configurations at FULL size on one B200 (bf16 tensor-core path): throughput, segment statistics and the
size-independent checks (bit-identical re-run, batch invariance of the first video).  Parity at these widths is tested
against the oracle at shortened T in tests/test_gpu_model.py; this script shows the full-size shapes run and how fast."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import config as C  # noqa: E402
from fact_clip_b200.models.blocks import FACT, FACT_CLIP  # noqa: E402
from fact_clip_b200.utils.synth import make_batch, make_text_embeddings  # noqa: E402

CASES = [
    ('config 1: gtea FACT, T=1024, C=11', 'gtea', 11, [1024] * 16, 33),
    ('config 2: breakfast FACT (MSTCN2, F=512), 4 ragged videos T~2000-6000, C=48', 'breakfast', 48, [2113, 3480, 4096, 5771], 8),
    ('config 3: havid holdout FACT_CLIP, T=4096, C=75', 'havid_view0_lh_pt_holdout', 75, [4096] * 16, 8),
    ('config 5 shape (inference only): epic FACT_CLIP (MSTCN2, M=300, fpos), T=16384, C=98', 'epic_shape', 98, [16384] * 2, 52),
]


def main():
    dev = 'cuda'
    for title, preset, ncls, lens, nseg in CASES:
        cfg = C.PRESETS[preset]()
        torch.manual_seed(0)
        clip = bool(cfg.use_clip)
        net = (FACT_CLIP(cfg, 2048, ncls, make_text_embeddings(ncls)) if clip else FACT(cfg, 2048, ncls)).eval()
        net.compute_mode = 'bf16'
        net = net.to(dev)
        xs, _ = make_batch(lens, 2048, ncls, base_seed=7, nseg=nseg)
        xd = [x.to(dev) for x in xs]
        a = net(xd, None)
        segs = [[int(v) for v in st['nseg'].tolist()] for st in net._last['blocks'] if 'nseg' in st]
        b = net(xd, None)
        same = all(np.array_equal(p['pred'], q['pred']) for p, q in zip(a, b))
        alone = net(xd[:1], None)
        inv = np.array_equal(alone[0]['pred'], a[0]['pred'])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pinned = [x.pin_memory() for x in xs]
        for _ in range(2):
            net.submit(pinned, None).result()
        e0.record()
        n = 5
        hs = [net.submit(pinned, None) for _ in range(n)]
        for h in hs:
            h.result()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps(dict(case=title, videos=len(lens), frames=sum(lens), ms_per_batch=round(ms, 3),
                              frames_per_s=round(sum(lens) / ms * 1e3), rerun_identical=bool(same), batch_invariant=bool(inv),
                              segments_per_U_block=[[min(s), max(s)] for s in segs],
                              classes_predicted=int(len(np.unique(np.concatenate([p['pred'] for p in a])))))), flush=True)
        del net
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
