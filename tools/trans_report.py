"""Per-video, per-tensor bf16 error report for the FACT.trans fixture (teacher-forced segmentation)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import fact_oracle as O  # noqa: E402
from conftest import load_golden  # noqa: E402
from test_gpu_model import build, compare_video  # noqa: E402
from fact_clip_b200 import config as C  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'tiny_m_iuU_trans'
g = load_golden(name)
net = build(g, 'bf16')
hp = O.hparams_from_cfg(C.tiny(**g['tiny_kwargs']), g['in_dim'], g['n_classes'])
vids = g['videos']
nU = sum(1 for b in hp['blocks'] if b['type'] == 'U')
forced = [[] for _ in range(nU)]
for v in vids:
    o = O.forward_video(g['state_dict'], hp, v['x'], clip=g['clip'], transcript=O.transcript_of(v['label']) if hp['trans'] else None)
    for u, p in enumerate([b['tdu_pred'] for b in o['blocks'] if 'tdu_pred' in b]):
        forced[u].append(p.cuda())
saves = net([v['x'].cuda() for v in vids], [v['label'].cuda() for v in vids], forced_preds=forced if nU else None)
for b, v in enumerate(vids):
    rep = []
    compare_video(net, b, v['blocks'], 10.0, report=rep)
    print('video', b, 'T', len(v['pred']), 'agree', float((saves[b]['pred'] == v['pred'].numpy()).mean()))
    for i, k, r in rep:
        print(f'   block {i} {k:16s} {r:.3e}')
