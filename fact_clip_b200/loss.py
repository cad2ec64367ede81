"""Training-loss VALUE of a batched forward, computed on the device (csrc/loss.cu).

Mirrors the reference's ``MatchCriterion`` (models/loss.py:47-277), ``infonce_contrastive_loss`` (loss.py:280-341) and the
``compute_loss`` / ``_loss_one_video`` methods of the model (models/blocks.py:90-106, 313-320, 369-382, 487-497, 677-786)
for all videos of a step at once.  What crosses PCIe per step: the ground-truth segment counts (B ints), the [M, S]
matching costs (the assignment itself is ``scipy.optimize.linear_sum_assignment`` on the host, like the reference) and the
final per-video numbers.  The reference instead moves every (1, T, M) attention map to the host and builds a (T, M, S)
numpy temporary per video (loss.py:91-106).

``LossRunner.seeds`` gives the gradient of the loss w.r.t. every logit tensor it reads (csrc/train_loss.cu): the entry points of
the hand-written backward pass of fact_clip_b200/train.py.
"""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

from . import ops

_TYPE = {'i': 0, 'u': 1, 'U': 2}
_TYPE_VN = {'I': 3, 'U': 4}          # verb/noun model (blocks_SepVerbNoun.py): halved frame / segment / token terms


class MatchCriterion:
    """Configuration holder with the reference's constructor (loss.py:49-53, built in scripts/train.py:207 as
    ``net.mcriterion = MatchCriterion(cfg, dataset.nclasses, dataset.bg_class)``)."""

    def __init__(self, cfg, nclasses, bg_ids=[], class_weight=None):
        self.cfg, self.nclasses, self.bg_ids, self._class_weight = cfg, nclasses, list(bg_ids), class_weight

    def class_weights(self):
        return class_weights(self)


def class_weights(crit):
    """cweight of set_label (loss.py:62-70): ones, null class last, background or per-class weights.  Reads only the
    attributes the reference's own MatchCriterion also has (cfg, nclasses, bg_ids, _class_weight), so the criterion the
    unchanged training script builds from ``fact_clip.models.loss`` (scripts/train.py:207) serves as well."""
    L = crit.cfg.Loss
    cw = torch.ones(crit.nclasses + 1)
    cw[-1] = float(L.nullw)
    if crit._class_weight is not None:
        cw[:crit.nclasses] = torch.as_tensor(crit._class_weight, dtype=torch.float32)[:crit.nclasses]
    else:
        for i in crit.bg_ids:
            cw[i] = float(L.bgw)
    return cw


def one_to_many(cost, transcript):
    """MatchCriterion._one_to_many_match (loss.py:155-194): Hungarian assignment of tokens to action CLASSES on the summed
    cost, leftover tokens to their cheapest class, then every ground-truth segment takes the cheapest token of its class.
    Returns (token index, segment index) lists grouped by ascending class, ascending segment inside a class."""
    actions, cls_of_seg = np.unique(transcript, return_inverse=True)
    t2a = np.stack([cost[:, cls_of_seg == j].sum(1) for j in range(len(actions))], axis=1)
    rows, cols = linear_sum_assignment(t2a)
    token_cls = np.full(cost.shape[0], -1, dtype=np.int64)
    token_cls[rows] = cols
    rest = token_cls < 0
    if rest.any():
        token_cls[rest] = t2a[rest].argmin(1)
    aind, sind = [], []
    for j in range(len(actions)):
        segs, toks = np.nonzero(cls_of_seg == j)[0], np.nonzero(token_cls == j)[0]
        assert len(toks), 'fewer tokens than action classes in the transcript'
        aind.extend(toks[cost[np.ix_(toks, segs)].argmin(0)].tolist())
        sind.extend(segs.tolist())
    return aind, sind


def assign(cost, transcript, mode):
    """cost (M, S) -> matched (token, segment) index lists (loss.py:119-153)."""
    M, S = cost.shape
    if mode == 'seq':
        assert M >= S, (M, S)
        return list(range(S)), list(range(S))
    if mode == 'o2o':
        a, s = linear_sum_assignment(cost)
        return a.tolist(), s.tolist()
    if mode == 'o2m':
        return one_to_many(cost, transcript)
    raise ValueError(mode)


class LossRunner:
    """Loss value of one engine forward (``FactEngine.run`` output) for the whole batch."""

    def __init__(self, engine, criterion):
        self.e, self.crit = engine, criterion

    def run(self, out, labels):
        e, cfg = self.e, self.crit.cfg
        hp, Lc = e.hp, cfg.Loss
        B, slot, ln, dev = e.B, e.slot, e.len, e.dev
        vn = e.vn is not None              # verb/noun model: the criterion sees the ACTION classes, the blocks' log-probs
        C, M, nb = (len(e.vn[0]) if vn else hp['n_classes']), e.ntok, len(hp['blocks'])
        assert nb <= 8, 'loss value: at most 8 blocks'
        assert C == self.crit.nclasses, f'criterion built for {self.crit.nclasses} classes, model has {C}'
        I32, buf = torch.int32, e.buf
        # ---- labels -> ground-truth segments (the TDU run-length kernel serves MatchCriterion.set_label, loss.py:56-60)
        label = buf('loss_label', (B, slot), I32)
        label.zero_()
        for b, y in enumerate(labels):
            label[b, :out['lengths'][b]].copy_(torch.as_tensor(y).to(I32), non_blocking=True)
        gseg, gstart = buf('loss_gseg', (B, slot), I32), buf('loss_gstart', (B, slot), I32)
        glen, gcen, gn = buf('loss_glen', (B, slot), I32), buf('loss_gcen', (B, slot), I32), buf('loss_gn', (B,), I32)
        ops.tdu_segment(label, gseg, gstart, glen, gcen, gn, len=ln)
        nseg = gn.cpu().numpy()                                    # host sync 1: B ints
        smax = int(nseg.max())
        cweight = class_weights(self.crit).to(dev)          # C+1 floats; the criterion's cfg may change between calls
        transcript, sweight = buf('loss_tr', (B, smax), I32), buf('loss_sw', (B, smax))
        # ---- InfoNCE bookkeeping (blocks.py:697-748): seen-class list, label remapping, per-class frame counts
        clip = 'projected_frame_embeddings' in out
        seen = cmap = inv_count = nvalid = None
        if clip:
            hold = set(cfg.holdout_classes) if 'holdout_classes' in cfg and cfg.holdout_classes else set()
            seen_l = [i for i in range(C) if i not in hold]
            seen = e.derived(('loss_seen', tuple(seen_l)), lambda: torch.tensor(seen_l, dtype=I32, device=dev))
            m = np.full(C, -1, dtype=np.int32)
            m[seen_l] = np.arange(len(seen_l), dtype=np.int32)
            cmap = e.derived(('loss_cmap', tuple(seen_l)), lambda: torch.from_numpy(m).to(dev))
            inv_count, nvalid = buf('loss_invc', (B, C)), buf('loss_nvalid', (B,), I32)
        ops.label_prep(label, gstart, gn, cweight, C, transcript, sweight, ln, cmap=cmap, inv_count=inv_count, nvalid=nvalid)
        # ---- matching on the LAST block (blocks.py:94-96): cost on device, assignment on the host
        last = out['blocks'][-1]
        Mp = last['a2f_attn_logit'].shape[2]
        overlap, cost = buf('loss_ov', (B, smax, Mp)), buf('loss_cost', (B, M, smax))
        if vn:         # cprob = exp(action_logp) (blocks_SepVerbNoun.py:100-101)
            ops.match_cost(last['a2f_attn_seg'], last['action_logp'], transcript, gstart, glen, gn, Lc.pc, Lc.a2fc, overlap, cost,
                           M, ridx=last['seg_label'], logp=True)
        elif 'a2f_attn_seg' in last:
            ops.match_cost(last['a2f_attn_seg'], last['action_clogit'], transcript, gstart, glen, gn, Lc.pc, Lc.a2fc, overlap, cost,
                           M, ridx=last['seg_label'])
        else:
            ops.match_cost(last['a2f_attn'], last['action_clogit'], transcript, gstart, glen, gn, Lc.pc, Lc.a2fc, overlap, cost, M)
        cost_h, tr_h = cost.cpu().numpy(), transcript.cpu().numpy()     # host sync 2: B x M x S floats
        kmax = max(smax, 1)
        aind_h, sind_h = np.zeros((B, kmax), np.int32), np.zeros((B, kmax), np.int32)
        inv_h, nm_h = np.full((B, smax), -1, np.int32), np.zeros(B, np.int32)
        matches = []
        for b in range(B):
            S = int(nseg[b])
            a, s = assign(cost_h[b, :, :S], tr_h[b, :S], Lc.match)
            # QUIRK kept from loss.py:221-222: sweight multiplies the k-th matched column, which needs K == S
            assert len(a) == S, f'video {b}: {len(a)} matched pairs for {S} segments (the reference cannot broadcast sweight)'
            aind_h[b, :S], sind_h[b, :S], nm_h[b] = a, s, S
            inv_h[b, s] = np.arange(S, dtype=np.int32)
            matches.append((a, s))
        up = lambda name, arr: buf(name, arr.shape, I32).copy_(torch.from_numpy(arr), non_blocking=True)
        aind, sind, inv, nm = up('loss_aind', aind_h), up('loss_sind', sind_h), up('loss_inv', inv_h), up('loss_nm', nm_h)
        # ---- per-block terms
        nchunk = ops.loss_nchunk(slot)
        ws = buf('loss_ws', (8 * nb + 3, B, nchunk))
        ws.zero_()
        npred = buf('loss_npred', (nb, B), I32)
        types = []
        zero_lse = e.zbuf('loss_zero_lse', (B, C), torch.float32) if vn else None      # log-prob inputs: nothing to subtract
        for i, (st, bc) in enumerate(zip(out['blocks'], hp['blocks'])):
            t0, ty = 8 * i, (_TYPE_VN if vn else _TYPE)[bc['type']]
            types.append(ty)
            if vn:
                # frame / segment losses and smoothing on the action log-probabilities (is_logit=False forms, loss.py:249-277),
                # the model's own token loss (blocks_SepVerbNoun.py:254-266)
                ridx, rlen, ns = st['seg_label'], st['seg_lens'], st['nseg']
                npred[i].copy_(ns)
                ops.loss_pick(st['frame_logp'], C, label, ws[t0 + 0], ln, w=cweight, col_lse=zero_lse)
                ops.loss_smooth(st['frame_logp'], C, ws[t0 + 1], ln, is_logp=True)
                ops.token_loss(st['action_logp'], aind, sind, nm, transcript, cweight, ws[t0 + 2], logp_mean=True)
                ops.loss_pick(st['seg_logp'], C, label, ws[t0 + 7], ln, w=cweight, ridx=ridx, rlen=rlen, col_lse=zero_lse)
                if ty == 4:
                    f2a, a2f = st['f2a_attn_logit'], st['a2f_attn_logit']
                    lse = buf('loss_collse', (B, f2a.shape[2]))
                    ops.col_lse(f2a, M, ns, lse)
                    ops.loss_pick(f2a, M, gseg, ws[t0 + 3], ln, cols=aind, ncols=nm, col_lse=lse, tmap=inv, w=sweight, ridx=ridx, rlen=rlen)
                    ops.loss_pick(a2f, M, gseg, ws[t0 + 4], ln, cols=aind, ncols=nm, tmap=inv, w=sweight, ridx=ridx, rlen=rlen)
                continue
            fc = st['frame_clogit']
            ops.loss_pick(fc, C, label, ws[t0 + 0], ln, w=cweight)                               # frame_loss, loss.py:249-261
            ops.loss_smooth(fc, C, ws[t0 + 1], ln)
            ops.token_loss(st['action_clogit'], aind, sind, nm, transcript, cweight, ws[t0 + 2])
            if ty == 0:
                continue
            f2a, a2f = st['f2a_attn_logit'], st['a2f_attn_logit']                              # rows: frames / segments
            lse = buf('loss_collse', (B, f2a.shape[2]))
            if ty == 1:
                ops.col_lse(f2a, M, ln, lse)
                ops.loss_pick(f2a, M, gseg, ws[t0 + 3], ln, cols=aind, ncols=nm, col_lse=lse, tmap=inv, w=sweight)
                ops.loss_pick(a2f, M, gseg, ws[t0 + 4], ln, cols=aind, ncols=nm, tmap=inv, w=sweight)
                ops.loss_smooth(f2a, M, ws[t0 + 5], ln)
                ops.loss_smooth(a2f, M, ws[t0 + 6], ln)
            else:
                ridx, rlen, ns = st['seg_label'], st['seg_lens'], st['nseg']
                npred[i].copy_(ns)
                ops.col_lse(f2a, M, ns, lse)
                ops.loss_pick(f2a, M, gseg, ws[t0 + 3], ln, cols=aind, ncols=nm, col_lse=lse, tmap=inv, w=sweight, ridx=ridx, rlen=rlen)
                ops.loss_pick(a2f, M, gseg, ws[t0 + 4], ln, cols=aind, ncols=nm, tmap=inv, w=sweight, ridx=ridx, rlen=rlen)
                ops.loss_pick(st['seg_clogit'], C, label, ws[t0 + 7], ln, w=cweight, ridx=ridx, rlen=rlen)   # frame_loss_tdu
        if clip:
            sim = out['clip_logit']                                                             # emb @ text^T / temp
            ops.loss_pick(sim, len(seen), label, ws[8 * nb + 0], ln, cols=seen, tmap=cmap)     # frame -> class
            lse = buf('loss_clslse', (B, C))
            ops.col_lse(sim, C, ln, lse, rmask0=label, rmap=cmap)
            ops.loss_pick(sim, len(seen), label, ws[8 * nb + 1], ln, cols=seen, tmap=cmap, col_lse=lse, w=inv_count)
        res = buf('loss_out', (B, 4 + nb))
        ops.loss_combine(ws, types, ln, npred, C, M, float(Lc.sw), res, use_clip=clip,
                         fact_w=float(cfg.CLIP.fact_loss_weight) if clip else 1.0,
                         con_w=float(cfg.CLIP.contrastive_weight) if clip else 0.0,
                         nseen=len(seen) if clip else 0, nvalid=nvalid)
        ctx = dict(label=label, gseg=gseg, gstart=gstart, glen=glen, gn=gn, transcript=transcript, tr_h=tr_h, aind=aind, sind=sind,
                   nm=nm, aind_h=aind_h, sind_h=sind_h, inv_h=inv_h, nseg_h=nseg, cweight=cweight, cmap=cmap, seen=seen,
                   nvalid=nvalid, smax=smax, types=types, clip=clip, lse_cls=lse if clip else None)
        return dict(values=res, matches=matches, ctx=ctx)

    def seeds(self, out, res, scale=1.0):
        """Gradient of ``scale * mean_b loss_b`` w.r.t. every logit tensor the loss reads, for the training step
        (fact_clip_b200/train.py): list of (Var, gradient).  Mirrors the formulas of ``run`` / factk_loss_combine term by term
        (csrc/train_loss.cu); the matching is a constant of the backward pass, like in the reference (loss.py:131 no_grad)."""
        e, cfg, c = self.e, self.crit.cfg, res['ctx']
        hp, Lc = e.hp, cfg.Loss
        B, slot, ln, dev = e.B, e.slot, e.len, e.dev
        C, M, nb = hp['n_classes'], e.ntok, len(hp['blocks'])
        F32 = torch.float32
        clip = c['clip']
        fact_w = float(cfg.CLIP.fact_loss_weight) if clip else 1.0
        con_w = float(cfg.CLIP.contrastive_weight) if clip else 0.0
        coef_f = np.full(B, scale * fact_w / (nb * B), np.float32)
        coef_c = np.full(B, scale * con_w / B, np.float32)
        if clip:                     # no frame with a seen label: the reference returns the un-weighted FACT loss (blocks.py:744-746)
            nv = c['nvalid'].cpu().numpy()
            coef_f[nv == 0] = scale / (nb * B)
            coef_c[nv == 0] = 0.0
        dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
        cf, cf_half, cc = dv(coef_f), dv(coef_f * 0.5), dv(coef_c)
        # per ground-truth segment: the matched token (column) and the weight at its match position; per token: multiplicity
        smax = max(c['smax'], 1)
        colmap_h, wmap_h = np.full((B, smax), -1, np.int32), np.zeros((B, smax), np.float32)
        mult_h = np.zeros((B, M), np.float32)
        cw_h = c['cweight'].cpu().numpy()
        for b in range(B):
            S = int(c['nseg_h'][b])
            for gs in range(S):
                k = int(c['inv_h'][b, gs])
                if k >= 0:
                    colmap_h[b, gs] = c['aind_h'][b, k]
                    wmap_h[b, gs] = cw_h[c['tr_h'][b, k]]       # QUIRK (loss.py:221-222): weight by match POSITION
            for k in range(S):
                mult_h[b, c['aind_h'][b, k]] += 1.0
        colmap, wmap, mult = dv(colmap_h), dv(wmap_h), dv(mult_h)
        z = lambda t: torch.zeros_like(t)
        seeds = []
        sw = float(Lc.sw)
        for i, (st, bc) in enumerate(zip(out['blocks'], hp['blocks'])):
            ty = c['types'][i]
            tdu = ty == 2
            fc, ac = st['frame_clogit_v'], st['action_clogit_v']
            gfc, gac = z(fc.v), z(ac.v)
            ops.loss_grad_ce_rows(fc.v, C, gfc, c['label'], c['cweight'], ln, cf_half if tdu else cf)
            if sw != 0.0:
                ops.loss_grad_smooth(fc.v, C, gfc, ln, cf, sw)
            ops.loss_grad_token(ac.v, gac, c['aind'], c['sind'], c['nm'], c['transcript'], c['cweight'], cf)
            seeds += [(fc, gfc), (ac, gac)]
            if ty == 0:
                continue
            f2a, a2f = st['f2a_logit_v'], st['a2f_logit_v']
            gf2a, ga2f = z(f2a.v), z(a2f.v)
            seg = dict(seg_label=st['seg_label'], seg_start=st['seg_start'], seg_len=st['seg_lens']) if tdu else None
            nrows = st['nseg'] if tdu else ln
            lse = torch.empty(B, f2a.v.shape[2], dtype=F32, device=dev)
            ops.col_lse(f2a.v, M, nrows, lse)
            ops.loss_grad_xattn(1, f2a.v, M, gf2a, c, colmap, wmap, smax, nrows, cf, seg=seg, col_lse=lse)
            ops.loss_grad_xattn(0, a2f.v, M, ga2f, c, colmap, wmap, smax, nrows, cf, seg=seg, mult=mult)
            if ty == 1 and sw != 0.0:
                ops.loss_grad_smooth(f2a.v, M, gf2a, ln, cf, sw)
                ops.loss_grad_smooth(a2f.v, M, ga2f, ln, cf, sw)
            seeds += [(f2a, gf2a), (a2f, ga2f)]
            if tdu:
                sc = st['seg_clogit_v']
                gsc = z(sc.v)
                ops.loss_grad_ce_rows(sc.v, C, gsc, c['label'], c['cweight'], nrows, cf_half, seg=seg)
                seeds.append((sc, gsc))
        if clip:
            sim = out['clip_logit_v']
            gs = z(sim.v)
            ops.loss_grad_infonce(sim.v, C, gs, c['label'], c['cmap'], c['lse_cls'], c['nvalid'], int(c['seen'].numel()), ln, cc)
            seeds.append((sim, gs))
        return seeds
