"""CPU restatement of the reference's training loss VALUE (no gradients) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product computes the
same numbers with the CUDA kernels of fact_clip_b200/csrc/loss.cu.

Restates, per video and in the reference's evaluation order,
  * MatchCriterion.set_label / match / _one_to_many_match / a2f_soft_iou        (models/loss.py:56-194)
  * action_token_loss, cross_attn_loss(_tdu), frame_loss(_tdu), smooth_loss      (models/loss.py:8-19, 196-277)
  * infonce_contrastive_loss                                                    (models/loss.py:280-341)
  * InputBlock / UpdateBlock / UpdateBlockTDU.compute_loss                      (models/blocks.py:313-320, 369-382, 487-497)
  * FACT._loss_one_video / FACT_CLIP._loss_one_video                            (models/blocks.py:90-106, 677-786)
on the per-block tensors :func:`fact_oracle.forward_video` returns.  Pinned by tests/golden/loss_*.pt, which hold the
loss values the UNMODIFIED reference computes in eval mode (tests/golden/make_loss_golden.py).

The reference quirks that change numbers are kept on purpose (each marked QUIRK below).
"""
import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

import fact_oracle as O


def loss_params(cfg, bg_ids=(), class_weight=None):
    """The cfg keys the loss reads (default.py:116-123, CLIP.*, holdout_classes) as a plain dict."""
    def g(node, k):
        return node[k] if isinstance(node, dict) else getattr(node, k)
    L = g(cfg, 'Loss')
    clip = g(cfg, 'CLIP') if 'CLIP' in cfg else None
    return dict(pc=float(g(L, 'pc')), a2fc=float(g(L, 'a2fc')), match=g(L, 'match'), bgw=float(g(L, 'bgw')),
                nullw=float(g(L, 'nullw')), sw=float(g(L, 'sw')), bg_ids=list(bg_ids), class_weight=class_weight,
                contrastive_weight=float(g(clip, 'contrastive_weight')) if clip is not None else 0.0,
                fact_loss_weight=float(g(clip, 'fact_loss_weight')) if clip is not None else 1.0,
                temp=float(g(clip, 'temp')) if clip is not None else 0.07,
                holdout=list(g(cfg, 'holdout_classes')) if 'holdout_classes' in cfg else [])


class Criterion:
    """MatchCriterion state after set_label (loss.py:56-87)."""

    def __init__(self, label, n_classes, lp):
        label = torch.as_tensor(label).long()
        seg_label, start, lens = O.run_length(label.numpy())
        self.label, self.C, self.lp = label, n_classes, lp
        self.seg_label = torch.from_numpy(seg_label)
        self.transcript = label[torch.from_numpy(start)]
        self.S = len(start)
        self.onehot_class = torch.zeros(len(label), n_classes)
        self.onehot_class[torch.arange(len(label)), label] = 1
        self.onehot_seg = torch.zeros(len(label), self.S)
        self.onehot_seg[torch.arange(len(label)), self.seg_label] = 1
        cw = torch.ones(n_classes + 1)
        cw[-1] = lp['nullw']
        sw = torch.ones(self.S)
        if lp['class_weight'] is not None:
            cw[:n_classes] = torch.as_tensor(lp['class_weight'], dtype=torch.float32)[:n_classes]
            sw = torch.as_tensor(lp['class_weight'], dtype=torch.float32)[self.transcript]
        else:
            for i in lp['bg_ids']:
                cw[i] = lp['bgw']
                sw[self.transcript == i] = lp['bgw']
        self.cweight, self.sweight = cw, sw


def soft_iou(a2f_attn, onehot_seg):
    """loss.py:91-106: a2f_attn (T,M) rows sum to 1, onehot_seg (T,S) -> iou (M,S).  Direct definition (the (T,M,S)
    temporary); the CUDA path uses the closed form union = colsum - overlap + |segment|."""
    a = a2f_attn.unsqueeze(-1).numpy()
    o = onehot_seg.unsqueeze(1).numpy()
    overlap = np.einsum('tax,txs->as', a, o)
    union = np.minimum(a + o, 1.0).sum(0)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.nan_to_num(overlap / union, nan=0.0)


def one_to_many(cost, transcript):
    """loss.py:155-194: tokens are first assigned to action CLASSES (Hungarian on the per-class summed cost, the
    leftover tokens to their cheapest class), then every ground-truth segment takes the cheapest token of its class."""
    tr = transcript.numpy()
    actions = np.unique(tr)
    t2a = np.stack([cost[:, tr == a].sum(1) for a in actions], axis=1)
    aid, cid = linear_sum_assignment(t2a)
    rest = [a for a in range(cost.shape[0]) if a not in aid]
    rest_c = t2a[rest].argmin(1) if rest else np.zeros(0, dtype=np.int64)
    all_aid = np.array(aid.tolist() + rest, dtype=np.int64)
    all_cid = np.array([actions[i] for i in cid.tolist() + rest_c.tolist()])
    token_class = np.zeros(cost.shape[0])
    token_class[all_aid] = all_cid
    match = {}
    for a in actions:
        segs = np.where(tr == a)[0]
        toks = np.where(token_class == a)[0]
        pick = cost[toks][:, segs].argmin(0)
        for s, k in zip(segs, pick):
            match[s] = toks[k]
    return list(match.values()), list(match.keys())


def match_tokens(crit, action_prob, a2f_attn):
    """loss.py:108-153.  action_prob (M,C+1) softmax of the LAST block's token logits, a2f_attn (T,M)."""
    lp = crit.lp
    if lp['match'] == 'seq':
        assert action_prob.shape[0] >= crit.S
        idx = torch.arange(crit.S)
        return idx, idx
    cost = np.zeros((action_prob.shape[0], crit.S), dtype=np.float32)
    if lp['pc'] > 0:
        cost = cost - lp['pc'] * action_prob.detach()[:, crit.transcript].numpy()
    if lp['a2fc'] > 0:
        cost = cost - lp['a2fc'] * soft_iou(a2f_attn.detach(), crit.onehot_seg)
    if lp['match'] == 'o2o':
        aind, sind = linear_sum_assignment(cost)
    else:
        aind, sind = one_to_many(cost, crit.transcript)
    return torch.as_tensor(np.asarray(aind), dtype=torch.int64), torch.as_tensor(np.asarray(sind), dtype=torch.int64)


def smooth_loss(logit):
    """loss.py:8-19 on (T,C) logits: mean over (T-1)*C of the clamped squared log-prob step (NaN when T == 1)."""
    lsm = torch.log_softmax(logit, dim=-1)
    return torch.clamp((lsm[1:] - lsm[:-1]) ** 2, min=0, max=16).mean()


def action_token_loss(crit, match, action_clogit):
    """loss.py:196-209: unmatched tokens are labelled with the null class; class-weighted cross entropy."""
    aind, sind = match
    A, C1 = action_clogit.shape
    clabel = torch.full((A,), C1 - 1, dtype=torch.long)
    clabel[aind] = crit.transcript[sind]
    logp = torch.log_softmax(action_clogit, -1)
    w = crit.cweight[clabel]
    return -(w * logp[torch.arange(A), clabel]).sum() / w.sum()


def cross_attn_loss(crit, match, attn_logit, dim):
    """loss.py:211-225.  attn_logit (T,M); dim=2: softmax over the matched tokens of a frame (a2f); dim=1: softmax over
    the frames of a matched token (f2a).  QUIRK: ``sweight`` (S,) multiplies the k-th matched COLUMN, i.e. it is indexed
    by the position in the match list, not by the segment the column is matched to."""
    aind, sind = match
    tgt = crit.onehot_seg[:, sind]
    logp = torch.log_softmax(attn_logit[:, aind], dim=dim - 1)
    return (-(logp * tgt) * crit.sweight).sum() / crit.onehot_seg.sum()


def zoom(onehot, seg_label, seg_lens):
    """Frame-level one-hot targets pooled to the predicted segments (loss.py:235-237, 258-261)."""
    z = torch.zeros(len(seg_lens), onehot.shape[1]).index_add_(0, seg_label, onehot)
    return z / seg_lens[:, None]


def cross_attn_loss_tdu(crit, match, attn_logit, seg_label, seg_lens, dim):
    """loss.py:227-247: the same with segment-level rows; targets are the fraction of a predicted segment inside each
    ground-truth segment (same QUIRK on sweight)."""
    aind, sind = match
    z = zoom(crit.onehot_seg, seg_label, seg_lens)
    logp = torch.log_softmax(attn_logit[:, aind], dim=dim - 1)
    return (-(logp * z[:, sind]) * crit.sweight).sum() / z.sum()


def frame_loss(crit, frame_clogit):
    """loss.py:249-261."""
    logp = torch.log_softmax(frame_clogit, -1)
    return (-(logp * crit.onehot_class) * crit.cweight[:frame_clogit.shape[-1]]).sum() / crit.onehot_class.sum()


def frame_loss_tdu(crit, seg_clogit, seg_label, seg_lens):
    """loss.py:263-277."""
    logp = torch.log_softmax(seg_clogit, -1)
    z = zoom(crit.onehot_class, seg_label, seg_lens)
    return (-(logp * z) * crit.cweight[:logp.shape[-1]]).sum() / z.sum()


def infonce(emb, text, labels, temp):
    """loss.py:280-341 with B = 1: mean of the frame->class cross entropy and the class->frame term (every class of
    ``text`` counts in the mean, a class without frames contributes 0)."""
    sim = emb @ text.t() / temp
    v2t = -torch.log_softmax(sim, -1)[torch.arange(len(labels)), labels].mean()
    tgt = torch.zeros(len(labels), text.shape[0])
    tgt[torch.arange(len(labels)), labels] = 1
    logp = torch.log_softmax(sim.t(), dim=1)
    t2v = (-(logp * tgt.t()).sum(1) / tgt.sum(0).clamp(min=1.0)).mean()
    return (v2t + t2v) / 2


def block_loss(crit, match, st, btype, sw):
    """InputBlock / UpdateBlock / UpdateBlockTDU.compute_loss (blocks.py:313-320, 369-382, 487-497).  ``st``: one entry of
    forward_video()['blocks']: f2a_attn_logit is (M,T) / (M,S) as the reference stores it, a2f_attn_logit (T,M) / (S,M)."""
    fl = frame_loss(crit, st['frame_clogit'])
    atk = action_token_loss(crit, match, st['action_clogit'])
    sm = smooth_loss(st['frame_clogit'])
    if btype == 'i':
        return fl + atk + sw * sm
    f2a_t, a2f = st['f2a_attn_logit'].t(), st['a2f_attn_logit']
    if btype == 'u':
        f2a = cross_attn_loss(crit, match, f2a_t, 1)
        a2f_l = cross_attn_loss(crit, match, a2f, 2)
        return atk + f2a + a2f_l + fl + sw * (smooth_loss(a2f) + smooth_loss(f2a_t) + sm)
    sl, ln = st['seg_label'], st['seg_lens'].float()
    seg = frame_loss_tdu(crit, st['seg_clogit'], sl, ln)
    f2a = cross_attn_loss_tdu(crit, match, f2a_t, sl, ln, 1)
    a2f_l = cross_attn_loss_tdu(crit, match, a2f, sl, ln, 2)
    return (fl + seg) / 2 + atk + f2a + a2f_l + sw * sm


def loss_video(out, hp, label, lp, text_embeddings=None):
    """FACT._loss_one_video (blocks.py:90-106) / FACT_CLIP._loss_one_video (blocks.py:677-786) on forward_video()'s
    output.  Returns a dict: loss, block_losses, match and, for the CLIP model, fact_loss / contrastive_loss."""
    crit = Criterion(label, hp['n_classes'], lp)
    last = out['blocks'][-1]
    match = match_tokens(crit, torch.softmax(last['action_clogit'], -1), last['a2f_attn'])
    losses = [block_loss(crit, match, st, bc['type'], lp['sw']) for st, bc in zip(out['blocks'], hp['blocks'])]
    fact = sum(losses) / len(losses)
    res = dict(loss=fact, block_losses=losses, match=match)
    if text_embeddings is not None and 'projected_frame_embeddings' in out:
        text, labels, emb = text_embeddings, crit.label, out['projected_frame_embeddings']
        if lp['holdout']:        # blocks.py:706-748: holdout classes leave the denominator; their frames leave the loss
            seen = torch.tensor([i for i in range(text.shape[0]) if i not in set(lp['holdout'])])
            mapper = torch.full((text.shape[0],), -1, dtype=torch.long)
            mapper[seen] = torch.arange(len(seen))
            text, labels = text[seen], mapper[crit.label]
            ok = labels != -1
            if not ok.all():
                if ok.sum() == 0:
                    return res
                labels, emb = labels[ok], emb[ok]
        con = infonce(emb, text, labels, lp['temp'])
        res.update(fact_loss=fact, contrastive_loss=con,
                   loss=lp['fact_loss_weight'] * fact + lp['contrastive_weight'] * con)
    return res
