"""Generate tests/golden/metrics_golden.json by running the UNMODIFIED reference metrics (fact_clip/utils/evaluate.py
Checkpoint.compute_metrics) on seeded synthetic ground truth / prediction pairs.  Run in the build container only:

    python tests/golden/make_metrics_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_yacs_shim'))
sys.path.insert(0, '/root/reference')

from fact_clip.utils.evaluate import Checkpoint, Video  # noqa: E402


def synth(rng, T, ncls, nseg):
    cuts = np.sort(rng.choice(np.arange(1, T), size=min(nseg - 1, T - 1), replace=False)) if T > 1 else np.zeros(0, dtype=np.int64)
    lab = np.zeros(T, dtype=np.int64)
    prev = 0
    for c in list(cuts) + [T]:
        lab[prev:c] = rng.integers(0, ncls)
        prev = c
    return lab


def corrupt(rng, gt, ncls, p_flip, jitter):
    pred = np.roll(gt, rng.integers(-jitter, jitter + 1))
    flip = rng.random(gt.size) < p_flip
    pred[flip] = rng.integers(0, ncls, flip.sum())
    # a few clean runs of a wrong class (over-segmentation)
    for _ in range(3):
        a = rng.integers(0, gt.size)
        pred[a:a + rng.integers(1, 40)] = rng.integers(0, ncls)
    return pred


def case(seed, n_videos, ncls, bg, holdout, seen, eval_edit=True, short_pred=False):
    rng = np.random.default_rng(seed)
    gts, preds = [], []
    for _ in range(n_videos):
        T = int(rng.integers(30, 900))
        gt = synth(rng, T, ncls, int(rng.integers(1, 25)))
        pred = corrupt(rng, gt, ncls, 0.02, 6)
        if short_pred:
            pred = pred[::int(rng.integers(2, 5))]      # prediction at a lower frame rate: expand_frame_label path
        gts.append(gt)
        preds.append(pred)
    ck = Checkpoint(0, bg_class=bg, eval_edit=eval_edit, holdout_classes=holdout, seen_classes=seen)
    ck.add_videos([Video('v%d' % i, gt_label=g, pred=p) for i, (g, p) in enumerate(zip(gts, preds))])
    m = ck.compute_metrics()
    return dict(seed=seed, bg=bg, holdout=holdout, seen=seen, eval_edit=eval_edit,
                gt=[g.tolist() for g in gts], pred=[p.tolist() for p in preds],
                metrics={k: float(v) for k, v in m.items()},
                per_class={str(k): v for k, v in ck.per_class_metrics.items()})


if __name__ == '__main__':
    cases = [
        case(1, 6, 11, [10], [], []),
        case(2, 5, 8, [], [], []),
        case(3, 6, 12, [0], [3, 7, 9], [1, 2, 4, 5, 6, 8, 10, 11]),
        case(4, 4, 6, [5], [], [], short_pred=True),
        case(5, 3, 4, [3], [], [], eval_edit=False),
    ]
    # degenerate: single-frame video, all background, prediction constant
    ck = Checkpoint(0, bg_class=[0])
    gts = [np.zeros(1, dtype=np.int64), np.array([0, 0, 1, 1, 1, 0]), np.array([2, 2, 2, 2])]
    preds = [np.zeros(1, dtype=np.int64), np.array([1, 1, 1, 1, 1, 1]), np.array([2, 2, 2, 2])]
    ck.add_videos([Video('v%d' % i, gt_label=g, pred=p) for i, (g, p) in enumerate(zip(gts, preds))])
    m = ck.compute_metrics()
    cases.append(dict(seed=-1, bg=[0], holdout=[], seen=[], eval_edit=True, gt=[g.tolist() for g in gts],
                      pred=[p.tolist() for p in preds], metrics={k: float(v) for k, v in m.items()},
                      per_class={str(k): v for k, v in ck.per_class_metrics.items()}))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'metrics_golden.json')
    with open(out, 'w') as f:
        json.dump(cases, f)
    print('wrote', out, [(c['seed'], {k: round(v, 3) for k, v in c['metrics'].items()}) for c in cases])
