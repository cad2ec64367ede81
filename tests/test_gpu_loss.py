"""Loss VALUE on the GPU (csrc/loss.cu through forward(compute_loss=True)) against (a) loss values computed by the unmodified
reference (tests/golden/loss_*.pt) and (b) the loss oracle on a BASELINE-config shape."""
import glob
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, GOLDEN, load_golden

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import fact_oracle as O  # noqa: E402
import loss_oracle as LO  # noqa: E402
from fact_clip_b200 import config as C  # noqa: E402
from fact_clip_b200.models.blocks import FACT, FACT_CLIP  # noqa: E402
from fact_clip_b200.models.loss import MatchCriterion  # noqa: E402
from fact_clip_b200.utils.synth import make_batch, make_text_embeddings  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = 'cuda'
LOSS_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, 'loss_*.pt')))
TOL = 1e-4      # relative, fp32 mode (the forward itself is within 1e-4 of the reference)


def close(a, b, tol=TOL):
    if math.isnan(b):
        return math.isnan(a)
    return abs(a - b) <= tol * max(1.0, abs(b))


@pytest.mark.parametrize('name', LOSS_CASES)
def test_loss_values_match_reference(name):
    lg = torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)
    g = load_golden(lg['fixture'])
    cfg = C.tiny(**g['tiny_kwargs'])
    cfg.merge(dict(Loss=lg['loss'], holdout_classes=list(lg['holdout'])))
    net = (FACT_CLIP(cfg, g['in_dim'], g['n_classes'], make_text_embeddings(g['n_classes'])) if g['clip']
           else FACT(cfg, g['in_dim'], g['n_classes']))
    net.load_state_dict(g['state_dict'], strict=False)
    net.compute_mode = 'fp32'
    net = net.to(DEV).eval()
    net.mcriterion = MatchCriterion(cfg, g['n_classes'], lg['bg_ids'])
    xs, ys = [v['x'].to(DEV) for v in g['videos']], [v['label'].to(DEV) for v in g['videos']]
    loss, saves = net(xs, ys, compute_loss=True)
    assert close(float(loss), lg['batch_loss'])
    for b, (sv, ref) in enumerate(zip(saves, lg['videos'])):
        assert [list(m) for m in net.last_match[b]] == ref['match'], f'video {b} match'
        assert close(sv['loss']['loss'], ref['loss']), (b, sv['loss'], ref['loss'])
        for a, r in zip(sv['block_losses'], ref['block_losses']):
            assert close(a, r), (b, sv['block_losses'], ref['block_losses'])
        if 'fact_loss' in ref:
            assert close(sv['loss']['fact_loss'], ref['fact_loss'])
            assert close(sv['loss']['contrastive_loss'], ref['contrastive_loss'])
        assert np.array_equal(sv['pred'], ref['pred'].numpy())
    # one video per call gives the same numbers as the batch (the kernels are batch-invariant)
    one, sv1 = net(xs[:1], ys[:1], compute_loss=True)
    assert close(float(one), lg['videos'][0]['loss'])


def test_loss_havid_shape_vs_oracle():
    """HAViD CLIP config (75 classes, 75 tokens, holdout classes, o2o matching), shortened T, fp32 mode, vs the oracle."""
    cfg = C.PRESETS['havid_view0_lh_pt_holdout']()
    cfg.merge(dict(Loss=dict(nullw=0.05, bgw=0.5)))
    ncls, lens = 75, [640, 333, 500]
    torch.manual_seed(0)
    net = FACT_CLIP(cfg, 2048, ncls, make_text_embeddings(ncls)).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    hp = O.hparams_from_cfg(cfg, 2048, ncls)
    lp = LO.loss_params(cfg, bg_ids=[0])
    xs, ys = make_batch(lens, 2048, ncls, base_seed=70, nseg=8)
    net.compute_mode = 'fp32'
    net = net.to(DEV)
    net.mcriterion = MatchCriterion(cfg, ncls, [0])
    loss, saves = net([x.to(DEV) for x in xs], [y.to(DEV) for y in ys], compute_loss=True)
    ref_tot = []
    for b, (x, y) in enumerate(zip(xs, ys)):
        with torch.no_grad():
            out = O.forward_video(sd, hp, x, clip=True, fast_gru=True)
            ref = LO.loss_video(out, hp, y, lp, text_embeddings=sd['text_embeddings'])
        if not np.array_equal(out['pred'].numpy(), saves[b]['pred']):
            continue            # a flipped argmax changes the segmentation, and with it every U-block term
        assert [list(m) for m in net.last_match[b]] == [m.tolist() for m in ref['match']]
        assert close(saves[b]['loss']['loss'], float(ref['loss']), 5e-4), (saves[b]['loss'], float(ref['loss']))
        assert close(saves[b]['loss']['contrastive_loss'], float(ref['contrastive_loss']), 5e-4)
        ref_tot.append(float(ref['loss']))
    assert len(ref_tot) >= 2


def test_loss_epic_shape_vs_oracle():
    """BASELINE config 5 shape (epic hyper-parameters on blocks.FACT_CLIP: MSTCN2, 300 tokens, fpos, one-to-many matching,
    nullw 0.05), shortened T, fp32 mode: loss value, InfoNCE term and match against the oracle."""
    cfg = C.PRESETS['epic_shape']()
    ncls, lens = 98, [900, 1100]
    torch.manual_seed(0)
    net = FACT_CLIP(cfg, 2048, ncls, make_text_embeddings(ncls)).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    hp = O.hparams_from_cfg(cfg, 2048, ncls)
    lp = LO.loss_params(cfg, bg_ids=[])
    xs, ys = make_batch(lens, 2048, ncls, base_seed=71, nseg=52)
    net.compute_mode = 'fp32'
    net = net.to(DEV)
    net.mcriterion = MatchCriterion(cfg, ncls, [])
    loss, saves = net([x.to(DEV) for x in xs], [y.to(DEV) for y in ys], compute_loss=True)
    checked = 0
    for b, (x, y) in enumerate(zip(xs, ys)):
        with torch.no_grad():
            out = O.forward_video(sd, hp, x, clip=True, fast_gru=True)
            ref = LO.loss_video(out, hp, y, lp, text_embeddings=sd['text_embeddings'])
        if not np.array_equal(out['pred'].numpy(), saves[b]['pred']):
            continue            # a flipped argmax changes the segmentation, and with it every U-block term
        assert [list(m) for m in net.last_match[b]] == [m.tolist() for m in ref['match']]
        assert close(saves[b]['loss']['loss'], float(ref['loss']), 5e-4), (saves[b]['loss'], float(ref['loss']))
        checked += 1
    assert checked >= 1
