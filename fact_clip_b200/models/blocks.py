"""FACT / FACT_CLIP drop-in modules (counterpart of fact_clip/models/blocks.py in the reference).

Same constructors, cfg keys, ``state_dict`` keys and ``forward(seq_list, label_list, compute_loss)``
outputs as the reference (FACT: blocks.py:19-135, FACT_CLIP: blocks.py:504-920); the forward itself
runs as hand-written sm_100a kernels through ``FactEngine`` -- batched over the whole ``seq_list``,
with no host synchronisation until the predictions are copied back.

Differences from the reference that a caller can observe:
  * ``seq_list`` is processed as ONE batch instead of a Python loop over videos (blocks.py:113-116).
  * the unused transcript (``torch_class_label_to_segment_label``, basic.py:38-54 -- a per-frame
    Python loop) is not computed when ``FACT.trans`` is False.
  * stashed attributes (``block.frame_clogit`` ...) are produced on request (``net.stash_video(i)``)
    rather than on every call; they refer to the LAST video of the batch by default, like the
    reference where each video overwrites the previous one.
"""
import os

import torch
import torch.nn as nn

from .. import config as cfgmod
from ..engine import FactEngine
from ..loss import LossRunner
from . import basic

try:  # the reference refuses to build FACT_CLIP without transformers (blocks.py:526-527); we do not need it
    import transformers  # noqa: F401
    CLIP_AVAILABLE = True
except ImportError:  # pragma: no cover
    CLIP_AVAILABLE = False


class FeatureProjection(nn.Module):
    """Linear -> LayerNorm -> ReLU -> Dropout -> Linear, then L2-normalise (blocks.py:141-175).
    Parameter holder; the math runs fused in the engine's CLIP head."""

    def __init__(self, feature_dim, clip_dim=512, hidden_dim=512, dropout=0.1):
        super().__init__()
        self.feature_dim, self.clip_dim = feature_dim, clip_dim
        self.projection = nn.Sequential(nn.Linear(feature_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(),
                                        nn.Dropout(dropout), nn.Linear(hidden_dim, clip_dim))


class Block(nn.Module):
    """Shared factories (blocks.py:204-240) and the attribute stash the loss / viz code reads."""

    def __str__(self):
        a2f = self.a2f_layer if hasattr(self, 'a2f_layer') else None
        f2a = self.f2a_layer if hasattr(self, 'f2a_layer') else None
        return (f'{type(self).__name__}(\n  f:{self.frame_branch},\n  a:{self.action_branch},\n'
                f'  a2f:{a2f},\n  f2a:{f2a}\n)')

    __repr__ = __str__

    @staticmethod
    def create_fbranch(cfg, in_dim=None, f_inmap=False):
        in_dim = cfg.f_dim if in_dim is None else in_dim
        if cfg.f == 'm':
            return basic.MSTCN(in_dim, cfg.f_dim, cfg.hid_dim, cfg.f_layers, dropout=cfg.dropout, ln=cfg.f_ln,
                               ngroup=cfg.f_ngp, in_map=f_inmap)
        if cfg.f == 'm2':
            return basic.MSTCN2(in_dim, cfg.f_dim, cfg.hid_dim, cfg.f_layers, dropout=cfg.dropout, ln=cfg.f_ln,
                                ngroup=cfg.f_ngp, in_map=f_inmap)
        raise ValueError(f"frame branch {cfg.f!r}: only 'm' (MSTCN) and 'm2' (MSTCN++) exist (blocks.py:208-213)")

    def create_abranch(self, cfg):
        if cfg.a == 'sa':
            layer = basic.SALayer(cfg.a_dim, cfg.a_nhead, dim_feedforward=cfg.a_ffdim, dropout=cfg.dropout, attn_dropout=cfg.dropout)
            return basic.SADecoder(cfg.a_dim, cfg.a_dim, cfg.hid_dim, layer, cfg.a_layers, in_map=False)
        if cfg.a == 'sca':
            layer = basic.SCALayer(cfg.a_dim, cfg.hid_dim, cfg.a_nhead, cfg.a_ffdim, dropout=cfg.dropout, attn_dropout=cfg.dropout)
            norm = nn.LayerNorm(cfg.a_dim)
            return basic.SCADecoder(cfg.a_dim, cfg.a_dim, cfg.hid_dim, layer, cfg.a_layers, norm=norm, in_map=False)
        if cfg.a in ('gru', 'gru_om'):      # GRU over the transcript tokens (blocks.py:225-228)
            assert self.cfg.FACT.trans
            return basic.ActionUpdate_GRU(cfg.a_dim, cfg.a_dim, cfg.hid_dim, cfg.a_layers, dropout=cfg.dropout,
                                          out_map=(cfg.a == 'gru_om'))
        raise ValueError(cfg.a)

    @staticmethod
    def create_cross_attention(cfg, outdim, kq_pos=True):
        return basic.X2Y_map(cfg.hid_dim, cfg.hid_dim, outdim, head_dim=cfg.hid_dim, dropout=cfg.dropout, kq_pos=kq_pos)


class InputBlock(Block):
    def __init__(self, cfg, in_dim, nclass):
        super().__init__()
        self.cfg, self.nclass = cfg, nclass
        self.frame_branch = self.create_fbranch(cfg.Bi, in_dim, f_inmap=True)
        self.action_branch = self.create_abranch(cfg.Bi)


class UpdateBlock(Block):
    def __init__(self, cfg, nclass):
        super().__init__()
        self.cfg, self.nclass = cfg, nclass
        c = cfg.Bu
        self.frame_branch = self.create_fbranch(c)
        self.f2a_layer = self.create_cross_attention(c, c.a_dim)
        self.action_branch = self.create_abranch(c)
        self.a2f_layer = self.create_cross_attention(c, c.f_dim)


class UpdateBlockTDU(Block):
    def __init__(self, cfg, nclass):
        super().__init__()
        self.cfg, self.nclass = cfg, nclass
        c = cfg.BU
        self.frame_branch = self.create_fbranch(c)
        self.seg_update = nn.GRU(c.hid_dim, c.hid_dim // 2, c.s_layers, bidirectional=True)
        self.seg_combine = nn.Linear(c.hid_dim, c.hid_dim)
        self.f2a_layer = self.create_cross_attention(c, c.a_dim)
        self.action_branch = self.create_abranch(c)
        self.a2f_layer = self.create_cross_attention(c, c.f_dim)
        self.sf_merge = nn.Sequential(nn.Linear(c.hid_dim + c.f_dim, c.f_dim), nn.ReLU())


class TDU:
    """Stand-in for basic.TemporalDownsampleUpsample holding what the loss reads (seg_label, seg_lens)."""

    def __init__(self, seg_label, seg_lens):
        self.seg_label, self.seg_lens, self.num_seg = seg_label, seg_lens, int(seg_lens.numel())


class _FactBase(nn.Module):
    """Everything FACT and FACT_CLIP share: constructor body, batched forward, stash, save_model."""

    def _build(self, cfg, in_dim, n_classes):
        self.cfg, self.num_classes, self.in_dim = cfg, n_classes, in_dim
        base = cfg.Bi
        self.frame_pe = basic.PositionalEncoding(base.hid_dim, max_len=10000, empty=(not cfg.FACT.fpos))
        self.channel_masking_dropout = nn.Dropout2d(p=cfg.FACT.cmr)
        if not cfg.FACT.trans:      # no transcript at training / inference: learned action queries (blocks.py:30-31)
            self.action_query = nn.Parameter(torch.randn([cfg.FACT.ntoken, 1, base.a_dim]))
        else:                       # transcript available: the tokens are its embedded actions (blocks.py:32-34)
            self.action_pe = basic.PositionalEncoding(base.a_dim, max_len=1000)
            self.action_embed = nn.Embedding(n_classes, base.a_dim)
        return base

    def _build_blocks(self, cfg, in_dim, n_classes, base):
        blocks = []
        for t in cfg.FACT.block:
            if t == 'i':
                blocks.append(InputBlock(cfg, in_dim, n_classes))
            elif t == 'u':
                cfgmod.update_from(cfg.Bu, base, inplace=True)
                base = cfg.Bu
                blocks.append(UpdateBlock(cfg, n_classes))
            elif t == 'U':
                cfgmod.update_from(cfg.BU, base, inplace=True)
                base = cfg.BU
                blocks.append(UpdateBlockTDU(cfg, n_classes))
            else:
                raise ValueError(f'FACT.block type {t!r} (the reference handles i/u/U only, blocks.py:38-48)')
        self.block_list = nn.ModuleList(blocks)
        self.mcriterion = None
        # 'bf16' (default) or 'fp32'; see DESIGN.md.  FACTK_MODE sets the default for scripts that only call the constructor
        self.compute_mode = os.environ.get('FACTK_MODE', 'bf16')
        assert self.compute_mode in ('bf16', 'fp32'), f'FACTK_MODE={self.compute_mode!r}'
        self._engine = None

    # -------------------------------------------------------------- engine plumbing
    def engine(self):
        if self._engine is None or self._engine.mode != self.compute_mode:
            hp = cfgmod.hparams(self.cfg, self.in_dim, self.num_classes)
            if hasattr(self, 'vids'):       # Epic verb/noun model: action id -> (verb id, noun id)
                hp['vn'] = (self.vids, self.nids)
            self._engine = FactEngine(self, hp, clip=isinstance(self, FACT_CLIP), mode=self.compute_mode)
        return self._engine

    def forward(self, seq_list, label_list=None, compute_loss=False, forced_preds=None):
        dev = next(self.parameters()).device
        if dev.type == 'cuda':
            with torch.cuda.device(dev):          # kernels launch on the current stream of the MODEL's device
                return self._forward(seq_list, label_list, compute_loss, forced_preds)
        return self._forward(seq_list, label_list, compute_loss, forced_preds)

    def _forward(self, seq_list, label_list=None, compute_loss=False, forced_preds=None):
        if compute_loss and self.mcriterion is None:
            raise RuntimeError('compute_loss=True needs net.mcriterion = MatchCriterion(cfg, nclasses, bg_ids) (scripts/train.py:207)')
        seqs = list(seq_list)        # CUDA tensors, or (pinned) host tensors copied straight into the packed batch
        if not seqs and not compute_loss:
            return []                # the reference's per-video loop simply does not run (blocks.py:113)
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('FACT forward runs only on a CUDA device through libfactk.so (no CPU fallback)')
        if any(s.dim() != 2 or s.shape[0] == 0 or s.shape[1] != self.in_dim for s in seqs):
            raise ValueError(f'every video must be a non-empty (T, {self.in_dim}) feature matrix, got '
                             f'{[tuple(s.shape) for s in seqs]}')
        if self.training:
            return self._forward_train(seqs, label_list, compute_loss, forced_preds)
        if self.cfg.FACT.trans:
            return self._forward_with_transcripts(seqs, label_list, forced_preds, compute_loss)
        # the verb/noun model's loss reads the action log-probabilities of EVERY block: keep them
        # the loss reads the cross-attention logits of EVERY block (and the verb/noun model its action log-probabilities): keep them
        keep = getattr(self, 'keep_attn', False) or compute_loss
        out = self.engine().run(seqs, forced_preds=forced_preds, keep=keep)
        self._last = out
        if compute_loss:
            res = LossRunner(self.engine(), self.mcriterion).run(out, label_list)
            self.last_match = res['matches']
        pred = out['pred'].cpu().numpy()            # the one D2H sync of the call (blocks.py:900)
        self._fire_branch_hooks(out)
        self.stash_video(len(seqs) - 1)
        saves = [{'pred': pred[b, :T].copy()} for b, T in enumerate(out['lengths'])]
        if not compute_loss:
            return saves
        # (final_loss, save_list) like blocks.py:120-130 / 902-914; the loss carries no graph (value only)
        vals = res['values'].cpu().numpy()
        nb = len(self.block_list)
        for b, s in enumerate(saves):
            s['loss'] = {'loss': float(vals[b, 0])}
            if vals[b, 3] > 0:
                s['loss'].update(fact_loss=float(vals[b, 1]), contrastive_loss=float(vals[b, 2]))
            s['block_losses'] = vals[b, 4:4 + nb].tolist()
        # attributes the training script reads after the call (scripts/train.py:286-289): those of the last video
        self.loss_list = [res['values'][-1, 4 + i] for i in range(nb)]
        if vals[-1, 3] > 0:
            self.fact_loss, self.contrastive_loss = res['values'][-1, 1], res['values'][-1, 2]
        return res['values'][:, 0].mean(), saves

    # -------------------------------------------------------------- training step
    def train_engine(self):
        from ..train import TrainEngine
        if getattr(self, '_train_engine', None) is None or self._train_engine.mode != self.compute_mode:
            hp = cfgmod.hparams(self.cfg, self.in_dim, self.num_classes)
            if hasattr(self, 'vids'):
                hp['vn'] = (self.vids, self.nids)
            self._train_engine = TrainEngine(self, hp, clip=isinstance(self, FACT_CLIP), mode=self.compute_mode)
        return self._train_engine

    def _forward_train(self, seqs, label_list, compute_loss, forced_preds=None):
        """``net.train(); loss, saves = net(seqs, labels, compute_loss=True); loss.backward()`` (scripts/train.py:262-264):
        train-mode forward (channel masking, time masking, dropouts: blocks.py:614-622) over the batch, loss on the device,
        and a loss tensor whose backward runs the hand-written gradient kernels (fact_clip_b200/train.py) and fills
        ``param.grad`` of every parameter.  ``net.grad_ready_hook(names, grads)``, when set, is called as soon as the gradients of
        a block's parameters are final (data-parallel all-reduce overlapped with the rest of the backward pass)."""
        eng = self.train_engine()
        eng.use_graphs = bool(getattr(self, 'train_graphs', False))      # CUDA-graph the step per batch shape (fixed-shape training)
        with torch.no_grad():
            out = eng.forward_train(seqs, forced_preds=forced_preds)
            self._last = out
            res = LossRunner(eng, self.mcriterion).run(out, label_list) if compute_loss else None
            pred = out['pred'].cpu().numpy()
        self._fire_branch_hooks(out)
        self.stash_video(len(seqs) - 1)
        saves = [{'pred': pred[b, :T].copy()} for b, T in enumerate(out['lengths'])]
        if not compute_loss:
            if isinstance(eng.tape, list):
                eng.tape = []
            return saves
        self.last_match = res['matches']
        vals = res['values'].cpu().numpy()
        nb = len(self.block_list)
        for b, s in enumerate(saves):
            s['loss'] = {'loss': float(vals[b, 0])}
            if vals[b, 3] > 0:
                s['loss'].update(fact_loss=float(vals[b, 1]), contrastive_loss=float(vals[b, 2]))
            s['block_losses'] = vals[b, 4:4 + nb].tolist()
        self.loss_list = [res['values'][-1, 4 + i] for i in range(nb)]
        if vals[-1, 3] > 0:
            self.fact_loss, self.contrastive_loss = res['values'][-1, 1], res['values'][-1, 2]
        names = [n for n, p in self.named_parameters() if p.requires_grad]
        params = [p for _, p in self.named_parameters() if p.requires_grad]
        loss = _TrainStep.apply(self, out, res, names, res['values'][:, 0].mean(), *params)
        return loss, saves

    def _forward_with_transcripts(self, seqs, label_list, forced_preds=None, compute_loss=False):
        """FACT.trans (blocks.py:74-79, 113-118): every video brings its own token count (the length of its transcript), so
        the videos run one per call like the reference; the transcript is the run-length coding of the label sequence
        (basic.py:38-54, vectorised: one unique_consecutive instead of a Python loop over the frames)."""
        assert label_list is not None and len(label_list) == len(seqs), 'FACT.trans needs the frame labels of every video'
        saves, dev = [], next(self.parameters()).device
        keep = getattr(self, 'keep_attn', False) or (compute_loss and hasattr(self, 'vids'))
        self._per_video = []
        total, matches = None, []
        for i, (seq, label) in enumerate(zip(seqs, label_list)):
            transcript = torch.unique_consecutive(torch.as_tensor(label).long()).to(dev)
            forced = None if forced_preds is None else [[u[i]] for u in forced_preds]
            out = self.engine().run([seq], forced_preds=forced, keep=keep, transcript=transcript)
            pred = out['pred'].cpu().numpy()
            saves.append({'pred': pred[0, :out['lengths'][0]].copy()})
            if compute_loss:      # the loss of this video (one token per transcript entry), blocks.py:120-126
                res = LossRunner(self.engine(), self.mcriterion).run(out, [label])
                vals = res['values'].cpu().numpy()
                nb = len(self.block_list)
                saves[-1]['loss'] = {'loss': float(vals[0, 0])}
                saves[-1]['block_losses'] = vals[0, 4:4 + nb].tolist()
                matches.append(res['matches'][0])
                total = res['values'][0, 0].clone() if total is None else total + res['values'][0, 0]
            if getattr(self, 'keep_attn', False):        # the engine's buffers are reused by the next video: keep copies for stash_video(i)
                self._per_video.append(_clone_tree(out))
        self._last = out
        self.stash_video(len(seqs) - 1)
        if compute_loss:
            self.last_match = matches
            return total / len(seqs), saves
        return saves

    def submit(self, seq_list, label_list=None, channel_major=False):
        """Asynchronous variant of ``forward`` for inference loops: enqueue the host->device copy (side stream,
        double-buffered), the kernels and the device->host copy of the predictions, and return a handle whose
        ``result()`` gives the same list ``forward`` returns.  Lets batch i+1's input copy overlap batch i's kernels."""
        if self.cfg.FACT.trans:
            raise NotImplementedError('FACT.trans models run one video per call through forward() (the transcript sets the token count)')
        if self.training:
            raise RuntimeError('the forward is built for eval mode: call net.eval() first (see forward())')
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('FACT forward runs only on a CUDA device through libfactk.so (no CPU fallback)')
        with torch.cuda.device(dev):
            h = self.engine().submit(list(seq_list), channel_major=channel_major)
        self._last = h.out
        return h

    def _fire_branch_hooks(self, out):
        """Forward hooks registered on ``block.frame_branch`` / ``block.action_branch`` (what
        scripts/fact_input_emb_logit_viz.py:24-30,62-63 does) fire once per video, in video order, with the tensor the
        reference's sub-module returns: the (T,1,H) / (M,1,H) feature BEFORE process_feature, i.e. the raw class logits in
        the class tail (the engine splices the probabilities in place, the logits are kept beside it).  The hook's
        ``input`` argument is an empty tuple: the engine never materialises per-module inputs."""
        blocks = [(blk, st) for blk, st in zip(self.block_list, out['blocks'])
                  if blk.frame_branch._forward_hooks or blk.action_branch._forward_hooks]
        if not blocks:
            return
        for b, T in enumerate(out['lengths']):
            for blk, st in blocks:
                for mod, feat, logit, rows in ((blk.frame_branch, st['frame_feature'], st['frame_clogit'], T),
                                               (blk.action_branch, st['action_feature'], st['action_clogit'], None)):
                    if not mod._forward_hooks:
                        continue
                    n = logit.shape[-1]
                    f, lg = feat[b, :rows].float(), logit[b, :rows, :n].float()
                    o = torch.cat([f[:, :f.shape[1] - n], lg], -1).unsqueeze(1)
                    for hook in list(mod._forward_hooks.values()):
                        hook(mod, (), o)

    def stash_video(self, b):
        """Expose video ``b`` of the last batch through the reference's per-block attributes
        (blocks.py:305-309, 359-366, 473-483) as views -- shapes (T,1,C), (M,1,C+1), (1,T,M)..."""
        out = self._last
        if self.cfg.FACT.trans:           # one engine call per video (see _forward_with_transcripts)
            if getattr(self, '_per_video', None):
                out, b = self._per_video[b], 0
            else:
                b = 0
        T, M = out['lengths'][b], int(out['blocks'][0]['action_clogit'].shape[1])
        for blk, st in zip(self.block_list, out['blocks']):
            C = self.num_classes
            blk.frame_clogit = st['frame_clogit'][b, :T].unsqueeze(1)
            blk.action_clogit = st['action_clogit'][b].unsqueeze(1)
            blk.action_feature = st['action_feature'][b, :, :-(C + 1)].unsqueeze(1)
            if 'nseg' in st:
                S = int(st['nseg'][b])
                lab = st['seg_label'][b, :T].long()
                blk.tdu = TDU(lab, st['seg_lens'][b, :S].long())
                blk.seg_clogit = st['seg_clogit'][b, :S].unsqueeze(1)
                # (the fused a2f kernel writes logits / attention to HBM only when they are read: net.keep_attn, the loss, and
                # the attention of the last block for the eval fusion)
                if st.get('a2f_attn_logit') is not None:
                    blk.a2f_attn_logit = st['a2f_attn_logit'][b, :S, :M].unsqueeze(0)
                if st.get('a2f_attn_seg') is not None:
                    blk.a2f_attn = st['a2f_attn_seg'][b, :S, :M][lab].unsqueeze(0)
                if st.get('f2a_attn_logit') is not None:      # (the fused f2a kernel never forms the logits unless keep_attn / the loss asks)
                    blk.f2a_attn_logit = st['f2a_attn_logit'][b, :S, :M].t().unsqueeze(0)
                if st.get('f2a_attn_seg') is not None:
                    blk.f2a_attn = st['f2a_attn_seg'][b, :S, :M][lab].t().unsqueeze(0)
            elif 'a2f_attn' in st:
                if st.get('a2f_attn_logit') is not None:
                    blk.a2f_attn_logit = st['a2f_attn_logit'][b, :T, :M].unsqueeze(0)
                if st.get('a2f_attn') is not None:
                    blk.a2f_attn = st['a2f_attn'][b, :T, :M].unsqueeze(0)
                if st.get('f2a_attn_logit') is not None:      # (the fused f2a kernel never forms the logits unless keep_attn / the loss asks)
                    blk.f2a_attn_logit = st['f2a_attn_logit'][b, :T, :M].t().unsqueeze(0)
                if st.get('f2a_attn') is not None:
                    blk.f2a_attn = st['f2a_attn'][b, :T, :M].t().unsqueeze(0)
        if 'projected_frame_embeddings' in out:
            self.projected_frame_embeddings = out['projected_frame_embeddings'][b, :T].unsqueeze(1)

    def save_model(self, fname):
        torch.save(self.state_dict(), fname)


class _TrainStep(torch.autograd.Function):
    """Autograd glue of the training step: the forward is already done (kernels, no graph); ``backward`` seeds the tape with
    the loss gradients and returns the parameter gradients the hand-written kernels produced."""

    @staticmethod
    def forward(ctx, net, out, res, names, value, *params):
        ctx.net, ctx.out, ctx.res, ctx.names, ctx.params = net, out, res, names, params
        return value.clone()

    @staticmethod
    def backward(ctx, gout):
        net = ctx.net
        eng = net._train_engine
        dev = next(net.parameters()).device
        with torch.no_grad(), torch.cuda.device(dev):
            seeds = LossRunner(eng, net.mcriterion).seeds(ctx.out, ctx.res, scale=float(gout))
            grads = eng.backward(seeds)
            # hand the gradient views (slices of the per-section flat buffers) to the parameters directly: no copy, and an
            # in-place all-reduce of a flat buffer is an all-reduce of the .grad tensors
            ret = []
            for n, p in zip(ctx.names, ctx.params):
                g = grads.get(n)
                if g is None or p.grad is not None:
                    ret.append(g)
                else:
                    p.grad = g
                    ret.append(None)
        ctx.out = ctx.res = None
        return (None, None, None, None, None) + tuple(ret)


def _clone_tree(o):
    if torch.is_tensor(o):
        return o.clone()
    if isinstance(o, dict):
        return {k: _clone_tree(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(_clone_tree(v) for v in o)
    return o


class FACT(_FactBase):
    def __init__(self, cfg, in_dim, n_classes):
        super().__init__()
        base = self._build(cfg, in_dim, n_classes)
        self._build_blocks(cfg, in_dim, n_classes, base)


class FACT_CLIP(_FactBase):
    """FACT + projection of frame features into CLIP text space + zero-shot logit head."""

    def __init__(self, cfg, in_dim, n_classes, text_embeddings=None):
        super().__init__()
        base = self._build(cfg, in_dim, n_classes)
        # blocks.py:568 computes this from Bi.hid_dim BEFORE the block loop; keep the order so
        # the RNG stream (and hence the random init) matches the reference constructor.
        self.frame_projection = FeatureProjection(feature_dim=base.hid_dim - n_classes, clip_dim=512,
                                                  hidden_dim=cfg.CLIP.projection_hidden_dim,
                                                  dropout=cfg.CLIP.projection_dropout)
        if text_embeddings is not None:
            self.register_buffer('text_embeddings', text_embeddings)
        else:
            self.text_embeddings = None
        self._build_blocks(cfg, in_dim, n_classes, base)
