"""CPU restatement of the Epic-Kitchens verb/noun model forward (reference models/blocks_SepVerbNoun.py) -- TEST
INFRASTRUCTURE ONLY (same rules as fact_oracle.py: only tests/, smoke() and the bench's cpu_baseline may import it).

Differences from blocks.py that this file follows line by line:
  * two class heads per feature (verbs | nouns), softmaxed separately (process_feature :229-234, logit2prob class_sep);
  * every action class a is the pair (VIDS[a], NIDS[a]); action (log-)probabilities are products (sums) of the pair's
    verb and noun (log-)probabilities (combine_verb_noun_to_action :188-226);
  * the segmentation of the temporal down-sampling is the argmax over the ACTION probabilities (:281-290);
  * the input block is down-sampled too: 2-layer bi-GRU over the segments, and the SCA decoder attends SEGMENTS
    (InputBlockTDU :356-398); the frame feature passes through the input block unchanged by the tokens;
  * evaluation works on the action log-probabilities (_eval :307-329).
Pinned by tests/golden/vn_*.pt generated from the unmodified reference (tests/golden/make_vn_golden.py).
"""
import numpy as np
import torch

import fact_oracle as O


def process_feature(x, n1, n2):
    """:229-234 -- the last n1+n2 channels are [verb | noun] logits, softmaxed separately."""
    clogit = x[:, -(n1 + n2):]
    prob = torch.cat([torch.softmax(clogit[:, :n1], -1), torch.softmax(clogit[:, n1:], -1)], -1)
    return torch.cat([x[:, :-(n1 + n2)], prob], -1), clogit


def combine(clogit, n1, vids, nids, action=False):
    """:188-226 with apply_log=True.  clogit (R, n1+n2) for frames / segments, (R, n1+1+n2+1) for tokens (action=True:
    both heads carry a null class, and the action vector gets a trailing null = verb null + noun null)."""
    k = n1 + 1 if action else n1
    v, n = torch.log_softmax(clogit[:, :k], -1), torch.log_softmax(clogit[:, k:], -1)
    a = v[:, vids] + n[:, nids]
    if action:
        a = torch.cat([a, (v[:, -1] + n[:, -1])[:, None]], -1)
    return a


def downsample(sd, p, frame, n1, n2, vids, nids, gru_layers, st, forced_pred=None):
    """temporal_downsample (:279-303): segments from the argmax of the action probabilities, segment mean, stacked
    bi-GRU, ReLU, seg_combine, process_feature."""
    cprob = frame[:, -(n1 + n2):]
    aprob = cprob[:, :n1][:, vids] * cprob[:, n1:][:, nids]
    pred = aprob.argmax(-1).numpy() if forced_pred is None else np.asarray(forced_pred)
    st['tdu_pred'] = torch.from_numpy(np.asarray(pred).astype(np.int64))
    seg_label, seg_start, seg_len = O.run_length(pred)
    sl, ln = torch.from_numpy(seg_label), torch.from_numpy(seg_len)
    st['seg_label'], st['seg_lens'] = sl, ln
    seg = torch.zeros(len(seg_len), frame.shape[1]).index_add_(0, sl, frame) / ln[:, None]
    for l in range(gru_layers):
        seg = O.gru_bidir(sd, p + 'seg_update.', seg, layer=l)
    seg = O.linear(torch.relu(seg), sd[p + 'seg_combine.weight'], sd[p + 'seg_combine.bias'])
    seg, st['seg_clogit'] = process_feature(seg, n1, n2)
    center = torch.from_numpy((seg_start + (seg_start + seg_len - 1)) // 2)
    return sl, seg, center


def input_block_tdu(sd, p, bc, n1, n2, vids, nids, frame, action, frame_pos, action_pos, st, forced_pred=None):
    """InputBlockTDU.forward (:367-398)."""
    frame = O.frame_branch(sd, p + 'frame_branch.', frame, bc, True)
    frame, st['frame_clogit'] = process_feature(frame, n1, n2)
    sl, seg, center = downsample(sd, p, frame, n1, n2, vids, nids, 2, st, forced_pred)
    seg_pos = frame_pos[center] if frame_pos is not None else None
    action = O.action_branch(sd, p + 'action_branch.', bc, action, action_pos, seg, seg_pos)
    action, st['action_clogit'] = process_feature(action, n1 + 1, n2 + 1)
    return frame, action


def update_block_tdu(sd, p, bc, n1, n2, vids, nids, frame, action, frame_pos, action_pos, st, s_layers, forced_pred=None):
    """UpdateBlockTDU.forward (:445-483)."""
    sl, seg, center = downsample(sd, p, frame, n1, n2, vids, nids, s_layers, st, forced_pred)
    seg_pos = frame_pos[center] if frame_pos is not None else None
    action, f2a_logit, f2a_attn = O.x2y_map(sd, p + 'f2a_layer.', seg, action, seg_pos, action_pos)
    action = O.action_branch(sd, p + 'action_branch.', bc, action, action_pos)
    action, st['action_clogit'] = process_feature(action, n1 + 1, n2 + 1)
    seg_out, a2f_logit, a2f_attn = O.x2y_map(sd, p + 'a2f_layer.', action, seg, action_pos, seg_pos)
    frame = torch.relu(O.linear(torch.cat([seg_out[sl], frame], -1), sd[p + 'sf_merge.0.weight'], sd[p + 'sf_merge.0.bias']))
    frame = O.frame_branch(sd, p + 'frame_branch.', frame, bc, False)
    frame, st['frame_clogit'] = process_feature(frame, n1, n2)
    st['f2a_attn_logit'], st['a2f_attn_logit'] = f2a_logit, a2f_logit
    st['f2a_attn'], st['a2f_attn'] = f2a_attn[:, sl], a2f_attn[sl]
    return frame, action


def evaluate(action_logp, a2f_attn, frame_logp, weight):
    """_eval (:307-329): as blocks.py's fusion, with exp(frame_logp) (NOT re-normalised) as the frame probability."""
    fprob = torch.exp(frame_logp)
    cpred = action_logp.argmax(1)
    loc = torch.where(cpred != action_logp.shape[-1] - 1)[0]
    if len(loc) == 0:
        return fprob.argmax(1)
    qprob = torch.exp(action_logp[:, :-1])
    qprob = qprob / qprob.sum(-1, keepdim=True)
    tok = loc[a2f_attn[:, loc].argmax(-1)]
    return ((1 - weight) * qprob[tok] + weight * fprob).argmax(1)


def forward_video(sd, hp, seq, vids, nids, forced_preds=None, transcript=None):
    """FACT._forward_one_video + eval (:58-93, 126, 337-341), eval mode.  hp: fact_oracle.hparams_from_cfg(...) with
    n_classes = (n1, n2); block string of 'I' / 'U'.  FACT.trans (:74-85): the tokens are cat[verb_embed(verb of the action),
    noun_embed(noun of the action)] + positional encoding, the query position is zero, and the prediction is the transcript
    entry whose token attends the frame most (_eval_w_transcript :331-336)."""
    n1, n2 = hp['n_classes']
    H = hp['blocks'][0]['hid_dim']
    T = seq.shape[0]
    frame_pos = O.positional_table(H, max(T, 1)) if hp['fpos'] else None
    vids, nids = torch.as_tensor(vids), torch.as_tensor(nids)
    if hp['trans']:
        assert transcript is not None
        A = 2 * sd['verb_embed.weight'].shape[1]
        action = torch.cat([sd['verb_embed.weight'][vids[transcript]], sd['noun_embed.weight'][nids[transcript]]], -1) \
            + O.positional_table(A, len(transcript))
        action_pos = torch.zeros_like(action)
    else:
        action_pos = sd['action_query'][:, 0]
        action = torch.zeros_like(action_pos)
    frame = seq
    out = {'blocks': []}
    for i, bc in enumerate(hp['blocks']):
        st, p = {}, f'block_list.{i}.'
        fp = None if forced_preds is None else forced_preds[i]
        if bc['type'] == 'I':
            frame, action = input_block_tdu(sd, p, bc, n1, n2, vids, nids, frame, action, frame_pos, action_pos, st, fp)
        else:
            frame, action = update_block_tdu(sd, p, bc, n1, n2, vids, nids, frame, action, frame_pos, action_pos, st,
                                             hp['s_layers'], fp)
        st['frame_logp'] = combine(st['frame_clogit'], n1, vids, nids)
        st['seg_logp'] = combine(st['seg_clogit'], n1, vids, nids)
        st['action_logp'] = combine(st['action_clogit'], n1, vids, nids, action=True)
        out['blocks'].append(st)
    last = out['blocks'][-1]
    if hp['trans']:
        out['pred'] = transcript[last['a2f_attn'][:, :len(transcript)].argmax(1)]
    else:
        out['pred'] = evaluate(last['action_logp'], last['a2f_attn'], last['frame_logp'], hp['mwt'])
    return out


# --------------------------------------------------------------------------------------------
# loss value (blocks_SepVerbNoun.py:95-112, 254-266, 400-413, 485-497 on loss.py's MatchCriterion)


def action_token_loss(crit, match, action_logp):
    """Block.action_token_loss (:254-266): unmatched tokens are labelled null, matched ones with their segment's action;
    class-weighted negative log-probability, MEAN over the tokens (no normalisation by the weights, unlike loss.py:196)."""
    aind, sind = match
    A, C1 = action_logp.shape
    clabel = torch.full((A,), C1 - 1, dtype=torch.long)
    clabel[aind] = crit.transcript[sind]
    return (-(action_logp[torch.arange(A), clabel]) * crit.cweight[clabel]).mean()


def loss_video(out, hp, label, lp, vids):
    """FACT._loss_one_video of the verb/noun model on forward_video()'s output: log-probability forms of the frame /
    segment losses (is_logit=False), (frame/2 + seg/2)/2 + token/2 (+ both cross-attention losses in update blocks) +
    sw * smooth, mean over the blocks.  The criterion sees the ACTION classes (len(vids) of them)."""
    import loss_oracle as LO
    A = len(vids)
    crit = LO.Criterion(label, A, lp)
    last = out['blocks'][-1]
    match = LO.match_tokens(crit, torch.exp(last['action_logp']), last['a2f_attn'])
    losses = []
    for st, bc in zip(out['blocks'], hp['blocks']):
        flp, slp = st['frame_logp'], st['seg_logp']
        sl, ln = st['seg_label'], st['seg_lens'].float()
        frame = (-(flp * crit.onehot_class) * crit.cweight[:A]).sum() / crit.onehot_class.sum() / 2
        z = LO.zoom(crit.onehot_class, sl, ln)
        seg = (-(slp * z) * crit.cweight[:A]).sum() / z.sum() / 2
        atk = action_token_loss(crit, match, st['action_logp']) / 2
        smooth = torch.clamp((flp[1:] - flp[:-1]) ** 2, min=0, max=16).mean()
        l = (frame + seg) / 2 + atk + lp['sw'] * smooth
        if bc['type'] == 'U':
            l = l + LO.cross_attn_loss_tdu(crit, match, st['f2a_attn_logit'].t(), sl, ln, 1) \
                  + LO.cross_attn_loss_tdu(crit, match, st['a2f_attn_logit'], sl, ln, 2)
        losses.append(l)
    return dict(loss=sum(losses) / len(losses), block_losses=losses, match=match)
