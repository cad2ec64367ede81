"""Inference sweep over synthetic videos, sharded over the ranks of one node (BASELINE config 4):

    python tools/sweep.py --videos 256 --batch 64
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py --videos 8192

Every rank owns a length-balanced shard of the video list (no data-path collective), runs it through the pipelined
public API (net.submit / result), the variable-length predictions are gathered on rank 0 (fact_clip_b200.parallel) and
scored there with the vectorised metrics (fact_clip_b200.metrics).  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fact_clip_b200 import config as C, metrics, parallel  # noqa: E402
from fact_clip_b200.models.blocks import FACT_CLIP  # noqa: E402
from fact_clip_b200.utils.synth import make_text_embeddings, make_video  # noqa: E402


def run(n_videos, batch, t_min, t_max, in_dim=2048, n_classes=75, preset='havid_view0_lh_pt_holdout', seed=0, pool=0):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)
    cfg = C.PRESETS[preset]()
    torch.manual_seed(0)
    net = FACT_CLIP(cfg, in_dim, n_classes, make_text_embeddings(n_classes)).eval().to(dev)
    lengths = np.random.default_rng(seed).integers(t_min, t_max + 1, n_videos).tolist()
    mine = parallel.shard_by_length(lengths, world)[rank]           # sorted by length: batches have similar slot sizes
    if pool:        # a sweep of thousands of videos: `pool` distinct synthetic videos per length, every video id maps to one of them
        made = {}
        def get(i):
            key = (lengths[i], i % pool)
            if key not in made:
                x, y = make_video(lengths[i], in_dim, n_classes, seed=seed * 100003 + key[1])
                made[key] = (x.pin_memory(), y)
            return made[key]
        vids = {i: get(i) for i in mine}
        host = {i: vids[i][0] for i in mine}
    else:
        vids = {i: make_video(lengths[i], in_dim, n_classes, seed=seed * 100003 + i) for i in mine}
        host = {i: vids[i][0].pin_memory() for i in mine}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pending, preds = [], {}
    for k in range(0, len(mine), batch):
        ids = mine[k:k + batch]
        pending.append((ids, net.submit([host[i] for i in ids], None)))
        if len(pending) > 1:
            ids0, h = pending.pop(0)
            preds.update({i: r['pred'] for i, r in zip(ids0, h.result())})
    for ids0, h in pending:
        preds.update({i: r['pred'] for i, r in zip(ids0, h.result())})
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tg = time.perf_counter()
    merged = parallel.gather_predictions(list(preds), [preds[i] for i in preds])       # the ONE collective of the sweep (NCCL)
    torch.cuda.synchronize()
    gather_s = time.perf_counter() - tg
    gts = parallel.gather_predictions(list(vids), [vids[i][1].numpy() for i in vids])
    out = None
    if rank == 0:
        order = sorted(merged)
        t1 = time.perf_counter()
        m = metrics.compute_metrics([gts[i] for i in order], [merged[i] for i in order], bg_class=[n_classes - 1],
                                    holdout_classes=list(cfg.holdout_classes) if getattr(cfg, 'holdout_classes', None) else [],
                                    seen_classes=[c for c in range(n_classes) if c not in set(getattr(cfg, 'holdout_classes', None) or [])])
        out = dict(videos=n_videos, frames=int(sum(lengths)), n_gpus=world, seconds=float(dt), frames_per_s=float(sum(lengths) / float(dt)),
                   gather_seconds=gather_s, gather_bytes=int(sum(lengths)) * 4, metrics_seconds=time.perf_counter() - t1, metrics={k: round(float(v), 4) for k, v in m.items()},
                   lengths=[int(min(lengths)), int(max(lengths))], batch=batch)
    if world > 1:
        dist.barrier()
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--videos', type=int, default=256)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--tmin', type=int, default=4096)
    ap.add_argument('--tmax', type=int, default=4096)
    ap.add_argument('--pool', type=int, default=0, help='distinct synthetic videos per rank (0: one per video id)')
    a = ap.parse_args()
    res = run(a.videos, a.batch, a.tmin, a.tmax, pool=a.pool)
    if res is not None:
        print(json.dumps(res))
