"""Golden fixtures of the Epic verb/noun model from the UNMODIFIED reference (models/blocks_SepVerbNoun.py); run in the
build container only:  python tests/golden/make_vn_golden.py

The reference reads the action -> (verb, noun) table from ./data/epic-kitchens/processed/*.txt relative to the working
directory (blocks_SepVerbNoun.py:147-170); the dataset is not shipped, so a small synthetic table is written to a scratch
directory and the reference is imported from there.  The table travels inside the fixture (vids / nids)."""
import os
import sys
import tempfile
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_yacs_shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings('ignore')

from fact_clip_b200 import config as ours  # noqa: E402
from fact_clip_b200.utils.synth import make_video  # noqa: E402
from make_golden import ref_cfg, SEED, SHARPEN  # noqa: E402

N1, N2 = 5, 7
PAIRS = [(0, 0), (1, 3), (4, 6), (2, 2), (1, 0), (3, 5), (0, 4), (4, 1), (2, 6)]       # action id -> (verb, noun)

CASES = [
    # name, tiny() kwargs, T list, in_dim
    ('vn_m2_IUU', dict(f='m2', block='IUU'), [88, 35, 1], 24),
    ('vn_m_IU_fpos', dict(f='m', block='IU', fpos=True, M=10), [60], 16),
    ('vn_m_IUU_trans', dict(f='m', block='IUU', trans=True), [70, 31, 2], 24),      # tokens = verb / noun embeddings of the transcript
]
# loss settings per case (Loss section, background action ids): one-to-one and one-to-many matching
LOSS = {'vn_m2_IUU': (dict(pc=0.5, a2fc=1.0, match='o2o', bgw=0.5, nullw=0.1, sw=2.0), [0]),
        'vn_m_IU_fpos': (dict(pc=1.0, a2fc=1.0, match='o2m', bgw=1.0, nullw=0.05, sw=5.0), []),
        'vn_m_IUU_trans': (dict(pc=1.0, a2fc=1.0, match='seq', bgw=1.0, nullw=0.1, sw=1.0), [])}


def write_tables(d):
    base = os.path.join(d, 'data', 'epic-kitchens', 'processed')
    os.makedirs(base)
    with open(os.path.join(base, 'verb_mapping.txt'), 'w') as f:
        f.write(''.join(f'{i} verb{i}\n' for i in range(N1)))
    with open(os.path.join(base, 'noun_mapping.txt'), 'w') as f:
        f.write(''.join(f'{i} noun{i}\n' for i in range(N2)))
    with open(os.path.join(base, 'mapping.txt'), 'w') as f:
        f.write(''.join(f'{a} verb{v},noun{n}\n' for a, (v, n) in enumerate(PAIRS)))


def main():
    scratch = tempfile.mkdtemp()
    write_tables(scratch)
    os.chdir(scratch)
    from fact_clip.models import blocks_SepVerbNoun as VN
    for name, kw, Ts, D in CASES:
        cfg = ref_cfg(ours.tiny(**kw))
        loss_kw, bg = LOSS[name]
        for k, v in loss_kw.items():
            cfg.Loss[k] = v
        torch.manual_seed(SEED)
        net = VN.FACT(cfg, D, N1, N2)
        from fact_clip.models.loss import MatchCriterion
        net.mcriterion = MatchCriterion(cfg, len(PAIRS), bg)
        assert VN._VIDS == [v for v, _ in PAIRS] and VN._NIDS == [n for _, n in PAIRS]
        net.eval()
        with torch.no_grad():
            for k, v in net.state_dict().items():
                if k.endswith(('out_linear.weight', 'conv_out.weight', 'seg_combine.weight')):
                    v.mul_(SHARPEN)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items() if not k.endswith('.pe')}
        vids = []
        for i, T in enumerate(Ts):
            x, y = make_video(T, D, len(PAIRS), seed=300 + i, nseg=min(5, T))
            with torch.no_grad():
                save = net([x], [y])
                loss, lsave = net([x], [y], compute_loss=True)
                crit = net.mcriterion
                crit.set_label(y)
                match = crit.match(torch.exp(net.block_list[-1].action_logp), net.block_list[-1].a2f_attn)
            loss_rec = dict(loss=float(loss), block_losses=[float(v) for v in net.loss_list], match=[m.tolist() for m in match])
            blocks = []
            for b in net.block_list:
                st = dict(frame_logp=b.frame_logp.detach().clone(), seg_logp=b.seg_logp.detach().clone(),
                          action_logp=b.action_logp.detach().clone(), seg_label=b.tdu.seg_label.clone(),
                          seg_lens=b.tdu.seg_lens.clone())
                for k in ('f2a_attn_logit', 'f2a_attn', 'a2f_attn_logit', 'a2f_attn'):
                    if getattr(b, k, None) is not None:
                        st[k] = getattr(b, k).detach().clone()
                blocks.append(st)
            vids.append(dict(x=x, label=y, pred=torch.from_numpy(save[0]['pred']), blocks=blocks, loss=loss_rec))
        torch.save(dict(name=name, tiny_kwargs=kw, n_classes=(N1, N2), vids=[v for v, _ in PAIRS], nids=[n for _, n in PAIRS],
                        in_dim=D, state_dict=sd, videos=vids, loss=loss_kw, bg_ids=bg), os.path.join(ROOT, 'tests', 'golden', name + '.pt'))
        print(name, 'ok', [[int(s['seg_lens'].numel()) for s in v['blocks']] for v in vids],
              [int(v['pred'].unique().numel()) for v in vids])


if __name__ == '__main__':
    main()
