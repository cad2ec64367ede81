"""The loss-value oracle (oracle/loss_oracle.py) against loss values computed by the unmodified reference
(tests/golden/loss_*.pt, written by tests/golden/make_loss_golden.py)."""
import glob
import math
import os
import sys

import pytest
import torch

from conftest import ROOT, GOLDEN, load_golden

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import fact_oracle as O  # noqa: E402
import loss_oracle as LO  # noqa: E402
from fact_clip_b200 import config as C  # noqa: E402

LOSS_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, 'loss_*.pt')))


def close(a, b, tol=2e-5):
    if math.isnan(b):
        return math.isnan(a)
    return abs(a - b) <= tol * max(1.0, abs(b))


def loss_cfg(lg, g):
    cfg = C.tiny(**g['tiny_kwargs'])
    cfg.merge(dict(Loss=lg['loss'], holdout_classes=list(lg['holdout'])))
    return cfg


@pytest.mark.parametrize('name', LOSS_CASES)
def test_loss_oracle_matches_reference_values(name):
    lg = torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)
    g = load_golden(lg['fixture'])
    cfg = loss_cfg(lg, g)
    hp = O.hparams_from_cfg(cfg, g['in_dim'], g['n_classes'])
    lp = LO.loss_params(cfg, bg_ids=lg['bg_ids'])
    text = g['state_dict'].get('text_embeddings') if g['clip'] else None
    totals = []
    for v, ref in zip(g['videos'], lg['videos']):
        with torch.no_grad():
            out = O.forward_video(g['state_dict'], hp, v['x'], clip=g['clip'],
                                  transcript=O.transcript_of(v['label']) if hp['trans'] else None)
            res = LO.loss_video(out, hp, v['label'], lp, text_embeddings=text)
        assert [m.tolist() for m in res['match']] == ref['match']
        for a, b in zip(res['block_losses'], ref['block_losses']):
            assert close(float(a), b), (float(a), b)
        assert close(float(res['loss']), ref['loss'])
        if 'fact_loss' in ref:
            assert close(float(res['fact_loss']), ref['fact_loss'])
            assert close(float(res['contrastive_loss']), ref['contrastive_loss'])
        totals.append(float(res['loss']))
    assert close(sum(totals) / len(totals), lg['batch_loss'])


def test_soft_iou_closed_form():
    """union = sum_t min(attn + onehot, 1) == colsum - overlap + |segment| when every attention value is <= 1: the form
    the CUDA cost kernel uses instead of the (T,M,S) temporary."""
    g = torch.Generator().manual_seed(3)
    T, M, S = 57, 9, 6
    attn = torch.softmax(torch.randn(T, M, generator=g) * 2, -1)
    seg = torch.sort(torch.randint(0, S, (T,), generator=g)).values
    onehot = torch.zeros(T, S)
    onehot[torch.arange(T), seg] = 1
    ref = LO.soft_iou(attn, onehot)
    overlap = onehot.t() @ attn                                  # (S,M)
    union = attn.sum(0)[None] - overlap + onehot.sum(0)[:, None]
    iou = torch.nan_to_num(overlap / union, nan=0.0).t()
    assert torch.allclose(iou, torch.from_numpy(ref), atol=1e-6)


def test_host_matching_equals_oracle_and_reference_semantics():
    """fact_clip_b200.loss.assign (the product's host-side Hungarian / one-to-many / sequential step on the device-computed
    costs) against the oracle's restatement of loss.py:119-194 on random cost matrices, repeated transcript classes included."""
    import numpy as np
    from fact_clip_b200 import loss as PL
    rng = np.random.default_rng(7)
    for M, S, ncls in ((12, 5, 3), (30, 9, 4), (8, 8, 8), (75, 11, 6), (5, 1, 1)):
        for _ in range(5):
            cost = rng.standard_normal((M, S)).astype(np.float32)
            tr = rng.integers(0, ncls, S)
            a, s = PL.assign(cost, tr, 'o2m')
            ra, rs = LO.one_to_many(cost, torch.from_numpy(tr))
            assert [int(x) for x in a] == [int(x) for x in ra] and [int(x) for x in s] == [int(x) for x in rs]
            assert sorted(s) == list(range(S))                       # every ground-truth segment is matched exactly once
            a, s = PL.assign(cost, tr, 'o2o')
            assert len(set(a)) == len(a) == S and sorted(s) == list(range(S))
            a, s = PL.assign(cost, tr, 'seq')
            assert a == list(range(S)) and s == list(range(S))
    with pytest.raises(ValueError):
        PL.assign(np.zeros((3, 2), np.float32), np.zeros(2, np.int64), 'nope')
