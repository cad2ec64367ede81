"""Epic verb/noun oracle (oracle/vn_oracle.py) against fixtures generated from the unmodified reference
(models/blocks_SepVerbNoun.py, tests/golden/make_vn_golden.py)."""
import glob
import os
import sys

import pytest
import torch

from conftest import ROOT, GOLDEN

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import fact_oracle as O  # noqa: E402
import vn_oracle as VO  # noqa: E402
from fact_clip_b200 import config as C  # noqa: E402

VN_CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, 'vn_*.pt')))


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize('name', VN_CASES)
def test_vn_oracle_matches_reference(name):
    g = torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)
    hp = O.hparams_from_cfg(C.tiny(**g['tiny_kwargs']), g['in_dim'], g['n_classes'])
    for v in g['videos']:
        with torch.no_grad():
            out = VO.forward_video(g['state_dict'], hp, v['x'], g['vids'], g['nids'],
                                   transcript=O.transcript_of(v['label']) if hp['trans'] else None)
        for st, ref in zip(out['blocks'], v['blocks']):
            assert torch.equal(st['seg_label'], ref['seg_label']) and torch.equal(st['seg_lens'], ref['seg_lens'])
            for k in ('frame_logp', 'seg_logp', 'action_logp'):
                assert rel(st[k], ref[k][:, 0]) < 1e-5, k
            for k in ('f2a_attn_logit', 'f2a_attn', 'a2f_attn_logit', 'a2f_attn'):
                if k in ref:
                    assert rel(st[k], ref[k][0]) < 1e-5, k
        assert torch.equal(out['pred'], v['pred'])


@pytest.mark.parametrize('name', VN_CASES)
def test_vn_loss_oracle_matches_reference(name):
    """Loss value of the verb/noun model (oracle/vn_oracle.loss_video) against the reference's compute_loss=True numbers."""
    import math
    import loss_oracle as LO
    g = torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)
    cfg = C.tiny(**g['tiny_kwargs'])
    cfg.merge(dict(Loss=g['loss']))
    hp = O.hparams_from_cfg(cfg, g['in_dim'], g['n_classes'])
    lp = LO.loss_params(cfg, bg_ids=g['bg_ids'])
    for v in g['videos']:
        with torch.no_grad():
            out = VO.forward_video(g['state_dict'], hp, v['x'], g['vids'], g['nids'],
                                   transcript=O.transcript_of(v['label']) if hp['trans'] else None)
            res = VO.loss_video(out, hp, v['label'], lp, g['vids'])
        ref = v['loss']
        assert [m.tolist() for m in res['match']] == ref['match']
        for a, b in zip(res['block_losses'], ref['block_losses']):
            assert (math.isnan(b) and math.isnan(float(a))) or abs(float(a) - b) <= 2e-5 * max(1.0, abs(b)), (float(a), b)
        b = ref['loss']
        assert (math.isnan(b) and math.isnan(float(res['loss']))) or abs(float(res['loss']) - b) <= 2e-5 * max(1.0, abs(b))
