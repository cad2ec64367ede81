"""cfg handling for the FACT / FACT_CLIP drop-in.

The reference drives its models with a yacs ``CfgNode`` (fact_clip/configs/default.py:1-154).  The
model constructors here accept a real yacs node OR this attribute-dict stand-in (yacs is not
installed in the build image) -- only the keys the forward path reads are required:
``FACT.{ntoken,block,trans,fpos,cmr,mwt}``, ``Bi/Bu/BU.*``, ``TM.*``, ``CLIP.*``, ``Loss.sw``.

Presets carry the hot-path values of the shipped YAMLs named in BASELINE.json's configs
(fact_clip/configs/{gtea,breakfast,havid_view0_lh_pt_holdout,epic-kitchens}.yaml).
"""
import copy


class CfgNode(dict):
    """Attribute-access dict with the small part of the yacs API the model path touches."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        return CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def defrost(self):
        return None

    def freeze(self):
        return None

    def merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge(v)
            else:
                self[k] = copy.deepcopy(v)
        return self

    def merge_from_file(self, fname):
        import yaml
        with open(fname) as f:
            return self.merge(CfgNode(yaml.safe_load(f)))


def update_from(cfg, ref, inplace=False):
    """Fill ``None`` fields of ``cfg`` from ``ref`` (configs/utils.py:219-231).  The reference calls
    this inside the model constructor with ``inplace=True``, mutating the caller's cfg; so do we."""
    if not inplace:
        cfg = cfg.clone()
    if hasattr(cfg, 'defrost'):
        cfg.defrost()
    for k in cfg:
        if k in ref and cfg[k] is None and ref[k] is not None:
            cfg[k] = ref[k]
    return cfg


_BLOCK_NONE = dict(hid_dim=None, dropout=None, a='sa', a_nhead=None, a_ffdim=None, a_layers=1, a_dim=None,
                   f=None, f_layers=5, f_ln=None, f_dim=None, f_ngp=None)


def defaults():
    """Model-relevant subset of default.py:44-148 (same default values)."""
    return CfgNode(dict(
        dataset='breakfast', split='split1', eval_bg=False, holdout_mode=False, holdout_classes=[],
        use_clip=False, batch_size=4,
        FACT=dict(ntoken=30, block='iuUU', trans=False, fpos=True, cmr=0.3, mwt=0.1),
        Bi=dict(hid_dim=512, dropout=0.5, a='sca', a_nhead=8, a_ffdim=2048, a_layers=6, a_dim=512,
                f='cnn', f_layers=10, f_ln=True, f_dim=512, f_ngp=4),
        Bu=dict(_BLOCK_NONE), BU=dict(_BLOCK_NONE, s_layers=1),
        Loss=dict(pc=1.0, a2fc=1.0, match='o2o', bgw=1.0, nullw=-1.0, sw=0.0),
        TM=dict(use=False, t=30, p=0.05, m=5, inplace=True),
        CLIP=dict(model_name='openai/clip-vit-base-patch32', text_trainable=True, temp=0.07,
                  precompute_text=True, use_prompt=True, text_emb_path=None, contrastive_weight=0.5,
                  fact_loss_weight=0.5, projection_hidden_dim=512, projection_dropout=0.1),
    ))


def _preset(bi, fact, extra=None):
    cfg = defaults()
    upd = dict(a_nhead=8, f_layers=10)
    cfg.merge(dict(Bi=dict(dict(a='sca', a_nhead=8, a_ffdim=512, a_layers=6, f_layers=10, f_ln=False,
                                f_ngp=1, hid_dim=512), **bi),
                   Bu=upd, BU=upd, FACT=fact,
                   Loss=dict(pc=0.2, sw=5.0)))
    if extra:
        cfg.merge(extra)
    return cfg


def gtea():
    """gtea.yaml: MSTCN, F=A=128, M=60, block iuU (BASELINE config 1; C=11)."""
    return _preset(dict(a_dim=128, f_dim=128, f='m', dropout=0.2),
                   dict(block='iuU', cmr=0.5, fpos=False, ntoken=60, mwt=0.1, trans=False),
                   dict(dataset='gtea', batch_size=1, TM=dict(use=True, t=60, p=0.1, m=5)))


def breakfast():
    """breakfast.yaml: MSTCN2, F=A=512, M=60, block iuUU (BASELINE config 2; C=48)."""
    return _preset(dict(a_dim=512, f_dim=512, f='m2', dropout=0.0),
                   dict(block='iuUU', cmr=0.3, fpos=False, ntoken=60, mwt=0.1, trans=False),
                   dict(dataset='breakfast', batch_size=4, eval_bg=True, TM=dict(use=True, t=30, p=0.05, m=5)))


def havid_view0_lh_pt_holdout():
    """havid_view0_lh_pt_holdout.yaml: MSTCN, F=A=256, M=75, iuUU, use_clip, temp 0.1 (BASELINE configs 3/4; C=75)."""
    return _preset(dict(a_dim=256, f_dim=256, f='m', dropout=0.2),
                   dict(block='iuUU', cmr=0.3, fpos=False, ntoken=75, mwt=0.1, trans=False),
                   dict(dataset='havid_view0_lh_pt', batch_size=2, eval_bg=True, use_clip=True, holdout_mode=True,
                        holdout_classes=[51, 53, 61, 67, 56], TM=dict(use=True, t=30, p=0.05, m=5),
                        CLIP=dict(temp=0.1)))


def epic_shape():
    """epic-kitchens.yaml hyper-parameters with block iUUU (the shipped 'IUUU' is not handled by
    blocks.py:38-48 -- SURVEY D1): MSTCN2, F=A=256, M=300, fpos (BASELINE config 5 shape)."""
    return _preset(dict(a_dim=256, f_dim=256, f='m2', dropout=0.0),
                   dict(block='iUUU', cmr=0.3, fpos=True, ntoken=300, mwt=0.1, trans=False),
                   dict(dataset='epic', batch_size=1, use_clip=True,
                        Loss=dict(match='o2m', nullw=0.05, bgw=0.5), TM=dict(use=False)))


def tiny(f='m', block='iuU', fpos=False, F=32, A=32, H=64, M=12, layers=4, a_layers=2, nhead=4, ffdim=48, trans=False,
         f_ln=False, f_ngp=1, a_i='sca', a_u='sa'):
    """A small configuration for golden fixtures and fast parity tests (not a shipped YAML)."""
    cfg = defaults()
    upd = dict(a_nhead=nhead, f_layers=layers, a=a_u)
    cfg.merge(dict(Bi=dict(a=a_i, a_dim=A, a_ffdim=ffdim, a_layers=a_layers, a_nhead=nhead, dropout=0.0,
                           f=f, f_dim=F, f_layers=layers, f_ln=f_ln, f_ngp=f_ngp, hid_dim=H),
                   Bu=upd, BU=upd,
                   FACT=dict(block=block, cmr=0.0, fpos=fpos, ntoken=M, mwt=0.1, trans=trans),
                   CLIP=dict(temp=0.1, projection_hidden_dim=40), use_clip=True))
    return cfg


PRESETS = dict(gtea=gtea, breakfast=breakfast, havid_view0_lh_pt_holdout=havid_view0_lh_pt_holdout,
               epic_shape=epic_shape)


_BLOCK_KEYS = ['hid_dim', 'dropout', 'a', 'a_nhead', 'a_ffdim', 'a_layers', 'a_dim', 'f', 'f_layers', 'f_ln', 'f_dim', 'f_ngp']


def hparams(cfg, in_dim, n_classes):
    """Flatten the cfg keys the forward reads (after the constructor's update_from) into plain dicts."""
    blocks = []
    for t in cfg.FACT.block:
        node = {'i': cfg.Bi, 'I': cfg.Bi, 'u': cfg.Bu, 'U': cfg.BU}[t]
        bc = {k: node[k] for k in _BLOCK_KEYS}
        bc['type'] = t
        blocks.append(bc)
    clip = cfg.CLIP if 'CLIP' in cfg else None
    return dict(in_dim=in_dim, n_classes=n_classes, blocks=blocks, ntoken=cfg.FACT.ntoken, fpos=bool(cfg.FACT.fpos),
                mwt=float(cfg.FACT.mwt), trans=bool(cfg.FACT.trans), s_layers=int(cfg.BU.s_layers),
                temp=float(clip.temp) if clip is not None else 0.07)
