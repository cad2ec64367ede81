"""Drop-in for the reference's Epic-Kitchens model (fact_clip/models/blocks_SepVerbNoun.py): separate verb and noun heads,
action classes as (verb, noun) pairs, temporal down-sampling in the input block as well.

Same constructor (``FACT(cfg, in_dim, n_classes1=98, n_classes2=301)``), parameter names, shapes and registration order
(so the random init under a seed and ``state_dict`` round trips match), same ``forward(seq_list, label_list)`` result.  The
action table comes from ``./data/epic-kitchens/processed/{verb_mapping,noun_mapping,mapping}.txt`` like the reference
(:147-170), or from the ``action_pairs`` argument (list of (verb id, noun id) per action id).  The forward runs on the GPU
through libfactk.so (fact_clip_b200/engine.py); ``compute_loss=True`` (eval mode) gives the reference's loss value with
``net.mcriterion = MatchCriterion(cfg, n_actions, bg_ids)``; FACT.trans models take their tokens from the verb / noun
embeddings of the transcript (:74-85) and run one video per call.
"""
import torch
import torch.nn as nn

from .. import config as cfgmod
from . import basic
from .blocks import Block, TDU, _FactBase


def load_action_pairs(base='./data/epic-kitchens/processed'):
    """(verb id, noun id) per action id from the three mapping files (blocks_SepVerbNoun.py:147-170; the line format of
    utils/dataset.py:23-35: '<id> <name>', and '<action id> <verb>,<noun>' in mapping.txt)."""
    def table(path):
        out = {}
        with open(path) as f:
            for line in f.read().split('\n')[:-1]:
                i, _, name = line.partition(' ')
                out[name] = int(i)
        return out
    v2i, n2i = table(f'{base}/verb_mapping.txt'), table(f'{base}/noun_mapping.txt')
    pairs = []
    with open(f'{base}/mapping.txt') as f:
        for line in f.read().split('\n')[:-1]:
            _, name = line.split(' ')
            v, n = name.split(',')
            pairs.append((v2i[v], n2i[n]))
    return pairs


class _VNBlock(Block):
    def __str__(self):
        return (f"{type(self).__name__}(\n  f:{self.frame_branch},\n  a:{self.action_branch},\n"
                f"  a2f:{getattr(self, 'a2f_layer', None)},\n  f2a:{getattr(self, 'f2a_layer', None)}\n)")


class InputBlockTDU(_VNBlock):
    def __init__(self, cfg, in_dim, nclass1, nclass2):
        super().__init__()
        self.cfg, self.nclass1, self.nclass2 = cfg, nclass1, nclass2
        c = cfg.Bi
        self.frame_branch = self.create_fbranch(c, in_dim, f_inmap=True)
        self.action_branch = self.create_abranch(c)
        self.seg_update = nn.GRU(c.hid_dim, c.hid_dim // 2, 2, bidirectional=True)
        self.seg_combine = nn.Linear(c.hid_dim, c.hid_dim)


class UpdateBlockTDU(_VNBlock):
    def __init__(self, cfg, nclass1, nclass2):
        super().__init__()
        self.cfg, self.nclass1, self.nclass2 = cfg, nclass1, nclass2
        c = cfg.BU
        self.frame_branch = self.create_fbranch(c)
        self.seg_update = nn.GRU(c.hid_dim, c.hid_dim // 2, c.s_layers, bidirectional=True)
        self.seg_combine = nn.Linear(c.hid_dim, c.hid_dim)
        self.f2a_layer = self.create_cross_attention(c, c.a_dim)
        self.action_branch = self.create_abranch(c)
        self.a2f_layer = self.create_cross_attention(c, c.f_dim)
        nn.Linear(c.hid_dim + c.f_dim, c.f_dim)      # :436 builds a Linear that :437 overwrites; it consumes the RNG stream
        self.sf_merge = nn.Sequential(nn.Linear(c.hid_dim + c.f_dim, c.f_dim), nn.ReLU())


class FACT(_FactBase):
    def __init__(self, cfg, in_dim, n_classes1=98, n_classes2=301, action_pairs=None):
        super().__init__()
        pairs = load_action_pairs() if action_pairs is None else [tuple(p) for p in action_pairs]
        self.vids, self.nids = [v for v, _ in pairs], [n for _, n in pairs]
        assert max(self.vids) + 1 == n_classes1 and max(self.nids) + 1 == n_classes2, \
            'the action table must use every verb / noun id range (blocks_SepVerbNoun.py:203-208)'
        self.cfg, self.in_dim = cfg, in_dim
        self.num_classes1, self.num_classes2 = n_classes1, n_classes2
        self.num_classes = (n_classes1, n_classes2)
        base = cfg.Bi
        self.frame_pe = basic.PositionalEncoding(base.hid_dim, max_len=10000, empty=(not cfg.FACT.fpos))
        self.channel_masking_dropout = nn.Dropout2d(p=cfg.FACT.cmr)
        if not cfg.FACT.trans:
            self.action_query = nn.Parameter(torch.randn([cfg.FACT.ntoken, 1, base.a_dim]))
        else:       # transcript tokens: halves from the verb and the noun of every action (:34-37)
            self.action_pe = basic.PositionalEncoding(base.a_dim, max_len=1000)
            self.verb_embed = nn.Embedding(n_classes1, base.a_dim // 2)
            self.noun_embed = nn.Embedding(n_classes2, base.a_dim // 2)
        blocks = []
        for t in cfg.FACT.block:
            if t == 'I':
                blocks.append(InputBlockTDU(cfg, in_dim, n_classes1, n_classes2))
            elif t == 'U':
                cfgmod.update_from(cfg.BU, base, inplace=True)
                base = cfg.BU
                blocks.append(UpdateBlockTDU(cfg, n_classes1, n_classes2))
            else:
                raise ValueError(t)
        self.block_list = nn.ModuleList(blocks)
        self.mcriterion = None
        self.compute_mode = 'bf16'
        self._engine = None

    def stash_video(self, b):
        """Per-block attributes of video ``b`` as the reference leaves them (:386-396, 470-481): frame_logp (T,1,A),
        seg_logp (S,1,A), action_logp (M,1,A+1), tdu, and the attention maps of the update blocks."""
        out = self._last
        if self.cfg.FACT.trans:           # one engine call per video (see _forward_with_transcripts)
            if getattr(self, '_per_video', None):
                out, b = self._per_video[b], 0
            else:
                b = 0
        T, M = out['lengths'][b], int(out['blocks'][0]['action_clogit'].shape[1])
        for blk, st in zip(self.block_list, out['blocks']):
            S = int(st['nseg'][b])
            lab = st['seg_label'][b, :T].long()
            blk.tdu = TDU(lab, st['seg_lens'][b, :S].long())
            if 'frame_logp' in st:
                blk.frame_logp = st['frame_logp'][b, :T].unsqueeze(1)
                blk.seg_logp = st['seg_logp'][b, :S].unsqueeze(1)
                blk.action_logp = st['action_logp'][b].unsqueeze(1)
            blk.f2a_attn = blk.a2f_attn = None
            if 'a2f_attn_logit' in st:
                blk.a2f_attn_logit = st['a2f_attn_logit'][b, :S, :M].unsqueeze(0)
                blk.a2f_attn = st['a2f_attn_seg'][b, :S, :M][lab].unsqueeze(0)
                if st.get('f2a_attn_logit') is not None:      # (the fused f2a kernel never forms the logits unless keep_attn / the loss asks)
                    blk.f2a_attn_logit = st['f2a_attn_logit'][b, :S, :M].t().unsqueeze(0)
                if st.get('f2a_attn_seg') is not None:
                    blk.f2a_attn = st['f2a_attn_seg'][b, :S, :M][lab].t().unsqueeze(0)
