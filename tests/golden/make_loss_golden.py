"""Loss-value fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_loss_golden.py

For fixtures written by make_golden.py: load the fixture's state_dict into the reference model, attach the reference's
MatchCriterion (scripts/train.py:207) and run ``net(seqs, labels, compute_loss=True)`` in eval mode (dropout, channel and
time masking are identities there, so the value is a deterministic function of the inputs).  Stores, per video, the total
loss, every block's loss (net.loss_list), the token<->segment match and the FACT / contrastive split of FACT_CLIP in
tests/golden/loss_<case>.pt -- numbers only, a few KB.
"""
import os
import sys
import warnings
import contextlib
import io

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
    ('o2m_holdout', 'tiny_m_iuU_clip', dict(pc=1.0, a2fc=1.0, match='o2m', bgw=1.0, nullw=0.1, sw=0.5), [], [2, 5]),
    ('o2o', 'tiny_m2_iuUU', dict(pc=0.2, a2fc=1.0, match='o2o', bgw=0.3, nullw=0.05, sw=5.0), [1], []),
    ('o2m', 'tiny_m2_iUU_fpos_clip', dict(pc=0.2, a2fc=1.0, match='o2m', bgw=1.0, nullw=0.05, sw=5.0), [0], []),
    ('seq', 'tiny_m_iu', dict(pc=0.0, a2fc=1.0, match='seq', bgw=1.0, nullw=0.2, sw=1.0), [], []),
    # transcript-conditioned model: one token per transcript entry, sequential matching (gtea_transcript.yaml style)
    ('trans_seq', 'tiny_m_iuU_trans', dict(pc=1.0, a2fc=1.0, match='seq', bgw=0.5, nullw=0.1, sw=2.0), [0], []),
    ('trans_o2o', 'tiny_m_iuU_trans', dict(pc=0.5, a2fc=1.0, match='o2o', bgw=1.0, nullw=0.1, sw=0.0), [], []),
]


def main():
    for variant, fixture, loss_kw, bg, holdout in CASES:
        g = torch.load(os.path.join(ROOT, 'tests', 'golden', fixture + '.pt'), weights_only=False)
        cfg = ref_cfg(ours.tiny(**g['tiny_kwargs']))
        for k, v in loss_kw.items():
            cfg.Loss[k] = v
        cfg.holdout_classes = list(holdout)
        C, D = g['n_classes'], g['in_dim']
        with contextlib.redirect_stdout(io.StringIO()):
            net = FACT_CLIP(cfg, D, C, make_text_embeddings(C)) if g['clip'] else FACT(cfg, D, C)
        missing = net.load_state_dict(g['state_dict'], strict=False)
        assert all(k.endswith('.pe') for k in missing.missing_keys) and not missing.unexpected_keys, missing
        net.eval()
        net.mcriterion = MatchCriterion(cfg, C, bg)
        vids = []
        for v in g['videos']:
            if hasattr(net, 'fact_loss'):
                del net.fact_loss, net.contrastive_loss
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                loss, saves = net([v['x']], [v['label']], compute_loss=True)
            crit = net.mcriterion
            last = net.block_list[-1]
            from fact_clip.models import basic
            match = crit.match(basic.logit2prob(last.action_clogit, dim=-1), last.a2f_attn)
            rec = dict(loss=float(loss), block_losses=[float(x) for x in net.loss_list], save_loss=saves[0]['loss'],
                       match=[m.tolist() for m in match], pred=torch.from_numpy(saves[0]['pred']))
            if hasattr(net, 'fact_loss'):
                rec.update(fact_loss=float(net.fact_loss), contrastive_loss=float(net.contrastive_loss))
            vids.append(rec)
        # the whole batch in one call: mean over the videos (blocks.py:130, 914)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            loss, _ = net([v['x'] for v in g['videos']], [v['label'] for v in g['videos']], compute_loss=True)
        name = f'loss_{fixture}_{variant}'
        torch.save(dict(name=name, fixture=fixture, loss=loss_kw, bg_ids=bg, holdout=holdout, videos=vids, batch_loss=float(loss)),
                   os.path.join(ROOT, 'tests', 'golden', name + '.pt'))
        print(name, 'batch', float(loss), [(r['loss'], len(r['match'][0])) for r in vids])


if __name__ == '__main__':
    main()
