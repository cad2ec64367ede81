"""Gradient fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_grad_golden.py

For forward fixtures written by make_golden.py: load the state_dict into the reference model, put it in TRAINING mode with
every random augmentation switched off in the cfg (dropout 0, FACT.cmr 0, TM.use False, CLIP.projection_dropout 0 -- then the
train-mode forward is a deterministic function of the inputs), run the reference's own training step
``loss, _ = net(seqs, labels, compute_loss=True); loss.backward()`` (scripts/train.py:262-264) on the whole batch and store the
loss and the gradient of EVERY parameter in tests/golden/grad_<case>.pt.
"""
import contextlib
import io
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_yacs_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings('ignore')

from fact_clip.models.blocks import FACT, FACT_CLIP  # noqa: E402
from fact_clip.models.loss import MatchCriterion  # noqa: E402
from fact_clip_b200 import config as ours  # noqa: E402
from fact_clip_b200.utils.synth import make_text_embeddings  # noqa: E402
from make_golden import ref_cfg  # noqa: E402

# name, forward fixture, Loss overrides, bg class ids, holdout classes
CASES = [
    ('o2o', 'tiny_m_iuU_clip', dict(pc=0.2, a2fc=1.0, match='o2o', bgw=0.5, nullw=0.05, sw=5.0), [0], []),
    ('o2m_holdout', 'tiny_m_iuU_clip_trained', dict(pc=1.0, a2fc=1.0, match='o2m', bgw=1.0, nullw=0.1, sw=0.5), [], [2, 5]),
    ('o2o', 'tiny_m2_iuUU_trained', dict(pc=0.2, a2fc=1.0, match='o2o', bgw=0.3, nullw=0.05, sw=5.0), [1], []),
    ('o2m', 'tiny_m2_iUU_fpos_clip', dict(pc=0.2, a2fc=1.0, match='o2m', bgw=1.0, nullw=0.05, sw=5.0), [0], []),
    ('seq', 'tiny_m_iu', dict(pc=0.0, a2fc=1.0, match='seq', bgw=1.0, nullw=0.2, sw=1.0), [], []),
    ('o2o', 'tiny_m_iuU_ln_ngp', dict(pc=0.2, a2fc=1.0, match='o2o', bgw=1.0, nullw=0.05, sw=2.0), [], []),
]


def main():
    for variant, fixture, loss_kw, bg, holdout in CASES:
        g = torch.load(os.path.join(ROOT, 'tests', 'golden', fixture + '.pt'), weights_only=False)
        cfg = ref_cfg(ours.tiny(**g['tiny_kwargs']))
        for k, v in loss_kw.items():
            cfg.Loss[k] = v
        cfg.holdout_classes = list(holdout)
        cfg.CLIP.projection_dropout = 0.0
        assert cfg.FACT.cmr == 0.0 and cfg.Bi.dropout == 0.0 and not cfg.TM.use
        C, D = g['n_classes'], g['in_dim']
        with contextlib.redirect_stdout(io.StringIO()):
            net = FACT_CLIP(cfg, D, C, make_text_embeddings(C)) if g['clip'] else FACT(cfg, D, C)
        net.load_state_dict(g['state_dict'], strict=False)
        net.train()
        net.mcriterion = MatchCriterion(cfg, C, bg)
        vids = [v for v in g['videos'] if v['x'].shape[0] > 1]           # a one-frame video has a NaN smoothing term
        with contextlib.redirect_stdout(io.StringIO()):
            loss, saves = net([v['x'] for v in vids], [v['label'] for v in vids], compute_loss=True)
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
        missing = [n for n, p in net.named_parameters() if p.grad is None]
        name = f'grad_{fixture}_{variant}'
        torch.save(dict(name=name, fixture=fixture, loss=loss_kw, bg_ids=bg, holdout=holdout, batch_loss=float(loss),
                        videos=[i for i, v in enumerate(g['videos']) if v['x'].shape[0] > 1], grads=grads, no_grad=missing),
                   os.path.join(ROOT, 'tests', 'golden', name + '.pt'))
        gn = torch.sqrt(sum((x.double() ** 2).sum() for x in grads.values()))
        print(name, 'loss', float(loss), 'params with grad', len(grads), 'without', missing, 'global grad norm', float(gn))


if __name__ == '__main__':
    main()
