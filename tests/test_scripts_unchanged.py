"""SURVEY section 4 "Script" row: the reference's own scripts run UNMODIFIED on the B200 kernels.

A synthetic mini-dataset in the reference's ``data/gtea`` layout (utils/dataset.py:181-189, 258-283) + a checkpoint are
written next to a scratch copy of the installed reference package (``get_project_base()`` is the package's parent, home.py:3-11);
``scripts/run_eval.py`` then runs twice as a subprocess -- once stock (the reference's eager PyTorch model on the GPU) and
once through ``python -m fact_clip_b200.dropin`` (same file, same arguments; ``fact_clip.models.blocks`` resolves to this
package) -- and the metrics the script writes must agree.  Skipped where no installed reference exists
(``baseline/_ref`` from tools/install_reference.py, or /root/reference in the build container).
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def _ref_root():
    for r in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(r, 'fact_clip')) and os.path.isdir(os.path.join(r, 'scripts')):
            return r
    return None


REF = _ref_root()
needs_ref = pytest.mark.skipif(REF is None, reason='no installed reference (baseline/_ref) to take the scripts from')

YAML = """
dataset: gtea
split: split1
batch_size: 2
use_clip: false
aux: {gpu: 0, debug: false}
FACT: {ntoken: 12, block: iuUU, trans: false, fpos: false, cmr: 0.0, mwt: 0.1}
Bi: {a: sca, a_dim: 64, a_ffdim: 48, a_layers: 2, a_nhead: 4, dropout: 0.0, f: m2, f_dim: 32, f_layers: 4, f_ln: false,
     f_ngp: 1, hid_dim: 64}
Bu: {a: sa, a_nhead: 4, f_layers: 4, a_layers: 1}
BU: {a: sa, a_nhead: 4, f_layers: 4, a_layers: 1, s_layers: 1}
"""


def make_project(tmp, g, n_videos=5, n_classes=None):
    """Scratch 'project': reference package + scripts, data/gtea/{mapping.txt, features/*.npy (D,T), groundTruth, splits}."""
    shutil.copytree(os.path.join(REF, 'fact_clip'), os.path.join(tmp, 'fact_clip'),
                    ignore=shutil.ignore_patterns('__pycache__'))
    shutil.copytree(os.path.join(REF, 'scripts'), os.path.join(tmp, 'scripts'))
    C, D = n_classes or g['n_classes'], g['in_dim']
    d = os.path.join(tmp, 'data', 'gtea')
    for sub in ('features', 'groundTruth', 'splits'):
        os.makedirs(os.path.join(d, sub))
    names = [f'c{i}' for i in range(C)]
    with open(os.path.join(d, 'mapping.txt'), 'w') as f:
        f.write(''.join(f'{i} {n}\n' for i, n in enumerate(names)))
    from fact_clip_b200.utils.synth import make_video
    vids = []
    for i in range(n_videos):
        x, y = (g['videos'][i]['x'], g['videos'][i]['label']) if i < len(g['videos']) else make_video(40 + 13 * i, D, g['n_classes'], seed=900 + i, nseg=5)
        v = f'vid{i}'
        np.save(os.path.join(d, 'features', v + '.npy'), x.numpy().T.copy())          # stored (D, T): the loader transposes
        with open(os.path.join(d, 'groundTruth', v + '.txt'), 'w') as f:
            f.write(''.join(names[int(c)] + '\n' for c in y))
        vids.append(v)
    for split in ('train', 'test'):
        with open(os.path.join(d, 'splits', f'{split}.split1.bundle'), 'w') as f:
            f.write(''.join(v + '.txt\n' for v in vids))
    with open(os.path.join(tmp, 'cfg.yaml'), 'w') as f:
        f.write(YAML)
    os.makedirs(os.path.join(tmp, 'run', 'ckpts'))
    ckpt = os.path.join(tmp, 'run', 'ckpts', 'network.iter-1.net')
    torch.save(g['state_dict'], ckpt)
    return ckpt


def run_script(tmp, script, args, dropin, mode='fp32'):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([tmp, os.path.join(ROOT, 'oracle', '_yacs_shim'), ROOT]),
               FACTK_MODE=mode, WANDB_MODE='disabled')
    cmd = [sys.executable] + (['-m', 'fact_clip_b200.dropin'] if dropin else []) + [os.path.join(tmp, 'scripts', script)] + args
    p = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, f'{" ".join(cmd)}\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}'
    return p.stdout


def read_metrics(tmp):
    """The script pickles its Checkpoint object (utils/evaluate.py:106-109); unpickle it where ``fact_clip`` is importable."""
    code = ('import gzip, pickle, json, sys; ck = pickle.load(gzip.open(sys.argv[1], "rb")); '
            'print(json.dumps({k: float(v) for k, v in ck.metrics.items()}))')
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([tmp, os.path.join(ROOT, 'oracle', '_yacs_shim')]))
    p = subprocess.run([sys.executable, '-c', code, os.path.join(tmp, 'run', 'eval_results', 'eval_result.gz')], env=env,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@needs_ref
def test_dropin_import_hook_resolves_model_modules(tmp_path):
    """``fact_clip.models.blocks`` -> this package; every other fact_clip module stays the reference's."""
    code = ('import fact_clip_b200.dropin, fact_clip.models.blocks as b, fact_clip.models.loss as l, fact_clip.utils.evaluate as e;'
            'import fact_clip_b200.models.blocks as ours;'
            'assert b.FACT is ours.FACT and b.FACT_CLIP is ours.FACT_CLIP, b.FACT;'
            'assert "fact_clip_b200" not in l.__file__ and "fact_clip_b200" not in e.__file__;'
            'print("ok")')
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REF, os.path.join(ROOT, 'oracle', '_yacs_shim'), ROOT]))
    p = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and 'ok' in p.stdout, p.stderr[-2000:]


@needs_ref
@pytest.mark.gpu
def test_run_eval_script_unmodified(tmp_path):
    tmp = str(tmp_path)
    g = load_golden('tiny_m2_iuUU_trained')
    ckpt = make_project(tmp, g)
    args = ['--cfg', os.path.join(tmp, 'cfg.yaml'), '--ckpt', ckpt]
    out_ref = run_script(tmp, 'run_eval.py', args, dropin=False)
    ck_ref = read_metrics(tmp)
    shutil.rmtree(os.path.join(tmp, 'run', 'eval_results'))
    out_new = run_script(tmp, 'run_eval.py', args, dropin=True)
    ck_new = read_metrics(tmp)
    m_ref, m_new = ck_ref, ck_new
    assert set(m_ref) == set(m_new) and {'Acc', 'Edit', 'F1@0.10', 'F1@0.25', 'F1@0.50'} <= set(m_new), m_new.keys()
    for k in m_ref:
        assert abs(m_ref[k] - m_new[k]) < 1e-6, (k, m_ref[k], m_new[k])
    assert m_new['Acc'] > 50, 'the trained fixture predicts most frames right; a collapsed prediction would not'


TRAIN_YAML = YAML + """
epoch: 12
lr: 0.001
optimizer: Adam
weight_decay: 0.0
clip_grad_norm: 10.0
Loss: {pc: 0.2, a2fc: 1.0, match: o2o, bgw: 1.0, nullw: 0.1, sw: 0.5}
TM: {use: true, t: 6, m: 2, p: 0.1}
"""


@needs_ref
@pytest.mark.gpu
def test_train_script_unmodified(tmp_path):
    """scripts/train.py as it is (wandb offline): 12 epochs x 3 batches of the synthetic mini-dataset through
    ``net(seqs, labels, compute_loss=True)`` -> ``loss.backward()`` -> ``clip_grad_norm_`` -> ``optimizer.step()``
    (scripts/train.py:262-268) on the hand-written training step, periodic evaluation and checkpointing included.  The run must
    finish, write its checkpoints and the FINISH_PROOF marker, and the training loss must have dropped."""
    tmp = str(tmp_path)
    g = load_golden('tiny_m2_iuUU_trained')
    make_project(tmp, g, n_videos=6, n_classes=11)      # the gtea loader hard-codes background class 10 (utils/dataset.py:189)
    with open(os.path.join(tmp, 'cfg_train.yaml'), 'w') as f:
        f.write(TRAIN_YAML.replace('cmr: 0.0', 'cmr: 0.2').replace('aux: {gpu: 0, debug: false}',
                                                                 'aux: {gpu: 0, debug: false, wandb_offline: true, print_every: 6, eval_every: 18}'))
    out = run_script(tmp, 'train.py', ['--cfg', os.path.join(tmp, 'cfg_train.yaml')], dropin=True)
    import glob
    import re
    logs = glob.glob(os.path.join(tmp, 'log', '**', 'FINISH_PROOF'), recursive=True)
    assert logs, out[-3000:]
    nets = glob.glob(os.path.join(os.path.dirname(logs[0]), 'ckpts', 'network.iter-*.net'))
    assert nets, 'no checkpoint written'
    sd = torch.load(nets[0], map_location='cpu')
    assert 'block_list.0.frame_branch.conv_out.weight' in sd and all(torch.isfinite(v).all() for v in sd.values() if v.is_floating_point())
    losses = [float(x) for x in re.findall(r'Iter\d+, loss:([0-9.]+)', out)]
    assert len(losses) >= 4 and losses[-1] < losses[0], losses
