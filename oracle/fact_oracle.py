"""CPU oracle for the FACT / FACT_CLIP forward hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.  The product
package (``fact_clip_b200``) never does; it fails loudly when its CUDA library is missing.

What it is: a plain, functional, fp32 restatement (torch CPU tensor arithmetic only -- no
``nn.Module``, no ``nn.GRU``, no ``nn.MultiheadAttention``) of the reference forward in
``/root/reference/fact_clip/models/blocks.py`` + ``basic.py``, driven directly by a reference-keyed
``state_dict``.  Every function cites the reference lines it follows.

Parity pinning: the reference ships NO tests / golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, imported in the build container
through ``oracle/_yacs_shim`` (script: ``tests/golden/make_golden.py``; fixtures committed under
``tests/golden/*.pt``; check: ``tests/test_oracle.py``).
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# hyper-parameters


def hparams_from_cfg(cfg, in_dim, n_classes):
    """Flatten the cfg keys the forward reads (blocks.py:21-50, 204-240, 510-608) into a dict.

    Applies the ``update_from`` inheritance of ``None`` fields (configs/utils.py:219-231) WITHOUT
    mutating ``cfg``.
    """
    def g(node, k):
        return node[k] if isinstance(node, dict) else getattr(node, k)

    keys = ['hid_dim', 'a', 'a_nhead', 'a_ffdim', 'a_layers', 'a_dim', 'f', 'f_layers', 'f_ln', 'f_dim', 'f_ngp']
    bi = {k: g(g(cfg, 'Bi'), k) for k in keys}
    blocks = []
    base = bi
    for t in g(g(cfg, 'FACT'), 'block'):
        if t in 'iI':
            cur = dict(bi)
        else:
            node = g(cfg, 'Bu' if t == 'u' else 'BU')
            cur = {k: (g(node, k) if g(node, k) is not None else base[k]) for k in keys}
            base = cur
        cur['type'] = t
        blocks.append(cur)
    clip = g(cfg, 'CLIP') if ('CLIP' in cfg) else None
    return dict(
        in_dim=in_dim, n_classes=n_classes, blocks=blocks,
        ntoken=g(g(cfg, 'FACT'), 'ntoken'), fpos=bool(g(g(cfg, 'FACT'), 'fpos')),
        mwt=float(g(g(cfg, 'FACT'), 'mwt')), trans=bool(g(g(cfg, 'FACT'), 'trans')),
        temp=float(g(clip, 'temp')) if clip is not None else 0.07,
        s_layers=int(g(g(cfg, 'BU'), 's_layers')),
    )


# --------------------------------------------------------------------------------------------
# small helpers


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def positional_table(d_model, length):
    """basic.py:90-102 -- sinusoid table, (length, d_model)."""
    pe = torch.zeros(length, d_model)
    position = torch.arange(0, length, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def add_pos(x, pos):
    """basic.py:313-320 -- add ``pos`` to the first ``pos.size(-1)`` channels only."""
    if pos is None:
        return x
    d = pos.shape[-1]
    x = x.clone()
    x[:, :d] = x[:, :d] + pos
    return x


def process_feature(x, nclass):
    """blocks.py:195-202 -- last ``nclass`` channels are logits; returns cat[feat, softmax], logits."""
    clogit = x[:, -nclass:]
    prob = torch.softmax(clogit, dim=-1)
    return torch.cat([x[:, :-nclass], prob], dim=-1), clogit


def conv1d_k3(x, w, b, dil):
    """nn.Conv1d(C, C, 3, padding=dil, dilation=dil, groups=g) on channels-last x (T, Cin); w (Cout, Cin/g, 3):
    output group j reads input group j only."""
    g = x.shape[1] // w.shape[1]
    if g > 1:
        ci, co = w.shape[1], w.shape[0] // g
        return torch.cat([conv1d_k3(x[:, j * ci:(j + 1) * ci], w[j * co:(j + 1) * co], b[j * co:(j + 1) * co], dil)
                          for j in range(g)], 1)
    T = x.shape[0]
    z = torch.zeros(dil, x.shape[1])
    xp = torch.cat([z, x, z], 0)
    y = b.unsqueeze(0).expand(T, -1).clone()
    for k in range(3):
        y = y + xp[k * dil:k * dil + T] @ w[:, :, k].t()
    return y


def conv1x1(x, w, b):
    return x @ w[:, :, 0].t() + b


# --------------------------------------------------------------------------------------------
# frame branch


def mstcn(sd, p, x, n_layers, in_map, ln=False):
    """basic.py:200-220 (+ DilatedResidualLayer.forward :154-171), eval mode; ``ln``: LayerNorm over the channels after
    the residual of every layer (f_ln, basic.py:166-169)."""
    if in_map:
        x = conv1x1(x, sd[p + 'conv_1x1.weight'], sd[p + 'conv_1x1.bias'])
    for i in range(n_layers):
        q = f'{p}layers.{i}.'
        h = torch.relu(conv1d_k3(x, sd[q + 'conv_dilated.weight'], sd[q + 'conv_dilated.bias'], 2 ** i))
        x = x + conv1x1(h, sd[q + 'conv_1x1.weight'], sd[q + 'conv_1x1.bias'])
        if ln:
            x = layer_norm(x, sd[q + 'norm.weight'], sd[q + 'norm.bias'])
    return conv1x1(x, sd[p + 'conv_out.weight'], sd[p + 'conv_out.bias'])


def mstcn2(sd, p, x, n_layers, in_map):
    """basic.py:263-281, eval mode."""
    if in_map:
        x = conv1x1(x, sd[p + 'conv_1x1_in.weight'], sd[p + 'conv_1x1_in.bias'])
    for i in range(n_layers):
        a = conv1d_k3(x, sd[f'{p}conv_dilated_1.{i}.weight'], sd[f'{p}conv_dilated_1.{i}.bias'], 2 ** (n_layers - 1 - i))
        b = conv1d_k3(x, sd[f'{p}conv_dilated_2.{i}.weight'], sd[f'{p}conv_dilated_2.{i}.bias'], 2 ** i)
        f = conv1x1(torch.cat([a, b], 1), sd[f'{p}conv_fusion.{i}.weight'], sd[f'{p}conv_fusion.{i}.bias'])
        x = torch.relu(f) + x
    return conv1x1(x, sd[p + 'conv_out.weight'], sd[p + 'conv_out.bias'])


def frame_branch(sd, p, x, bc, in_map):
    if bc['f'] == 'm':
        return mstcn(sd, p, x, bc['f_layers'], in_map, ln=bool(bc['f_ln']))
    assert not bc['f_ln'], 'MSTCN++ has no layer norm (basic.py:226)'
    return mstcn2(sd, p, x, bc['f_layers'], in_map)


# --------------------------------------------------------------------------------------------
# attention


def mha(sd, p, query, key, value, nhead):
    """torch.nn.MultiheadAttention forward (seq-first, B=1, eval) restated.

    Packed ``in_proj_weight`` when kdim == embed_dim, else separate q/k/v matrices.
    Returns (out (L,E), attn (nhead, L, S)).
    """
    E = query.shape[-1]
    bias = sd[p + 'in_proj_bias']
    if (p + 'in_proj_weight') in sd:
        W = sd[p + 'in_proj_weight']
        wq, wk, wv = W[:E], W[E:2 * E], W[2 * E:]
    else:
        wq, wk, wv = sd[p + 'q_proj_weight'], sd[p + 'k_proj_weight'], sd[p + 'v_proj_weight']
    q = linear(query, wq, bias[:E])
    k = linear(key, wk, bias[E:2 * E])
    v = linear(value, wv, bias[2 * E:])
    dh = E // nhead
    L, S = q.shape[0], k.shape[0]
    q = q.view(L, nhead, dh).transpose(0, 1)
    k = k.view(S, nhead, dh).transpose(0, 1)
    v = v.view(S, nhead, dh).transpose(0, 1)
    attn = torch.softmax((q @ k.transpose(1, 2)) / math.sqrt(dh), dim=-1)
    o = (attn @ v).transpose(0, 1).reshape(L, E)
    return linear(o, sd[p + 'out_proj.weight'], sd[p + 'out_proj.bias']), attn


def sca_decoder(sd, p, tgt, memory, pos, query_pos, bc):
    """basic.py:542-557 (SCADecoder) over basic.py:494-523 (SCALayer), post-norm, eval."""
    nh = bc['a_nhead']
    for i in range(bc['a_layers']):
        q = f'{p}layers.{i}.'
        qk = add_pos(tgt, query_pos)
        t2, _ = mha(sd, q + 'self_attn.', qk, qk, tgt, nh)
        tgt = layer_norm(tgt + t2, sd[q + 'norm1.weight'], sd[q + 'norm1.bias'])
        t2, _ = mha(sd, q + 'multihead_attn.', add_pos(tgt, query_pos), add_pos(memory, pos), memory, nh)
        tgt = layer_norm(tgt + t2, sd[q + 'norm2.weight'], sd[q + 'norm2.bias'])
        t2 = linear(torch.relu(linear(tgt, sd[q + 'linear1.weight'], sd[q + 'linear1.bias'])),
                    sd[q + 'linear2.weight'], sd[q + 'linear2.bias'])
        tgt = layer_norm(tgt + t2, sd[q + 'norm3.weight'], sd[q + 'norm3.bias'])
    tgt = layer_norm(tgt, sd[p + 'norm.weight'], sd[p + 'norm.bias'])          # blocks.py:223-224
    return linear(tgt, sd[p + 'out_linear.weight'], sd[p + 'out_linear.bias'])


def sa_decoder(sd, p, x, pos, bc):
    """basic.py:578-593 (SADecoder) over basic.py:429-452 (SALayer): q=k=x+pos, v=x."""
    nh = bc['a_nhead']
    for i in range(bc['a_layers']):
        q = f'{p}layers.{i}.'
        xp = add_pos(x, pos)
        t2, _ = mha(sd, q + 'multihead_attn.', xp, xp, x, nh)
        x = layer_norm(x + t2, sd[q + 'norm1.weight'], sd[q + 'norm1.bias'])
        t2 = linear(torch.relu(linear(x, sd[q + 'linear1.weight'], sd[q + 'linear1.bias'])),
                    sd[q + 'linear2.weight'], sd[q + 'linear2.bias'])
        x = layer_norm(x + t2, sd[q + 'norm2.weight'], sd[q + 'norm2.bias'])
    return linear(x, sd[p + 'out_linear.weight'], sd[p + 'out_linear.bias'])


def x2y_map(sd, p, X, Y, X_pos, Y_pos):
    """basic.py:349-389 with kq_pos=True: K from X+Xpos, V from X, Q from Y+Ypos, 1 head.

    Returns (Y_out, attn_logit (Y,X), attn (Y,X)).
    """
    xk = linear(add_pos(X, X_pos), sd[p + 'X_K.weight'], sd[p + 'X_K.bias'])
    xv = linear(X, sd[p + 'X_V.weight'], sd[p + 'X_V.bias'])
    yq = linear(add_pos(Y, Y_pos), sd[p + 'Y_Q.weight'], sd[p + 'Y_Q.bias'])
    logit = (yq @ xk.t()) / math.sqrt(xk.shape[-1])
    attn = torch.softmax(logit, dim=-1)
    feat = attn @ xv
    out = linear(torch.cat([Y, feat], -1), sd[p + 'Y_W.weight'], sd[p + 'Y_W.bias'])
    return out, logit, attn


# --------------------------------------------------------------------------------------------
# temporal down/up-sampling


def run_length(pred):
    """utils/utils.py:25-48 + basic.py:597-607: maximal runs -> (seg_label (T,), seg_start, seg_len)."""
    pred = np.asarray(pred)
    T = len(pred)
    flag = np.ones(T, dtype=bool)
    flag[1:] = pred[1:] != pred[:-1]
    start = np.nonzero(flag)[0]
    lens = np.diff(np.append(start, T))
    seg_label = np.cumsum(flag) - 1
    return seg_label.astype(np.int64), start.astype(np.int64), lens.astype(np.int64)


def gru_bidir(sd, p, x, layer=0):
    """One layer of nn.GRU(H, H/2, n, bidirectional=True), h0 = 0, seq-first, B=1 (blocks.py:401,432)."""
    S = x.shape[0]
    outs = []
    for suffix, order in (('', range(S)), ('_reverse', range(S - 1, -1, -1))):
        w_ih, w_hh = sd[f'{p}weight_ih_l{layer}{suffix}'], sd[f'{p}weight_hh_l{layer}{suffix}']
        b_ih, b_hh = sd[f'{p}bias_ih_l{layer}{suffix}'], sd[f'{p}bias_hh_l{layer}{suffix}']
        Hh = w_hh.shape[1]
        gi = x @ w_ih.t() + b_ih                    # (S, 3Hh) gate order r, z, n
        h = torch.zeros(Hh)
        out = torch.zeros(S, Hh)
        for t in order:
            gh = w_hh @ h + b_hh
            r = torch.sigmoid(gi[t, :Hh] + gh[:Hh])
            z = torch.sigmoid(gi[t, Hh:2 * Hh] + gh[Hh:2 * Hh])
            n = torch.tanh(gi[t, 2 * Hh:] + r * gh[2 * Hh:])
            h = (1 - z) * n + z * h
            out[t] = h
        outs.append(out)
    return torch.cat(outs, -1)


def action_gru(sd, p, x, bc):
    """basic.py:283-308 (ActionUpdate_GRU): stacked bidirectional GRU over the tokens (eval: no inter-layer dropout),
    LayerNorm, optional out_map ('gru_om')."""
    for l in range(bc['a_layers']):
        x = gru_bidir(sd, p + 'gru.', x, layer=l)
    x = layer_norm(x, sd[p + 'layernorm.weight'], sd[p + 'layernorm.bias'])
    if bc['a'] == 'gru_om':
        x = linear(x, sd[p + 'out_map.weight'], sd[p + 'out_map.bias'])
    return x


def action_branch(sd, p, bc, action, action_pos, frame=None, frame_pos=None):
    """blocks.py:215-232: 'sca' (input block), 'sa' (update blocks), 'gru' / 'gru_om' (either; ignores the memory)."""
    if bc['a'] in ('gru', 'gru_om'):
        return action_gru(sd, p, action, bc)
    if bc['a'] == 'sca':
        return sca_decoder(sd, p, action, frame, frame_pos, action_pos, bc)
    assert bc['a'] == 'sa', bc['a']
    return sa_decoder(sd, p, action, action_pos, bc)


def gru_bidir_fast(sd, p, x):
    """Same function through aten's fused CPU GRU (what the reference's nn.GRU calls); used for the
    timed cpu_baseline only.  tests/test_oracle.py checks it against :func:`gru_bidir`."""
    flat = [sd[f'{p}{n}_l0{s}'] for s in ('', '_reverse') for n in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh')]
    Hh = flat[1].shape[1]
    out, _ = torch._VF.gru(x.unsqueeze(1), torch.zeros(2, 1, Hh), flat, True, 1, 0.0, False, True, False)
    return out[:, 0]


# --------------------------------------------------------------------------------------------
# blocks


def input_block(sd, p, bc, C, frame, action, frame_pos, action_pos, st):
    """blocks.py:295-311."""
    frame = frame_branch(sd, p + 'frame_branch.', frame, bc, True)
    frame, st['frame_clogit'] = process_feature(frame, C)
    action = action_branch(sd, p + 'action_branch.', bc, action, action_pos, frame, frame_pos)
    action, st['action_clogit'] = process_feature(action, C + 1)
    return frame, action


def update_block(sd, p, bc, C, frame, action, frame_pos, action_pos, st):
    """blocks.py:343-367."""
    action, st['f2a_attn_logit'], st['f2a_attn'] = x2y_map(sd, p + 'f2a_layer.', frame, action, frame_pos, action_pos)
    action = action_branch(sd, p + 'action_branch.', bc, action, action_pos)
    action, st['action_clogit'] = process_feature(action, C + 1)
    frame, st['a2f_attn_logit'], st['a2f_attn'] = x2y_map(sd, p + 'a2f_layer.', action, frame, action_pos, frame_pos)
    frame = frame_branch(sd, p + 'frame_branch.', frame, bc, False)
    frame, st['frame_clogit'] = process_feature(frame, C)
    return frame, action


def update_block_tdu(sd, p, bc, C, frame, action, frame_pos, action_pos, st, forced_pred=None, fast_gru=False):
    """blocks.py:417-485.  ``forced_pred`` (T,) teacher-forces the segmentation."""
    pred = frame[:, -C:].argmax(-1).numpy() if forced_pred is None else np.asarray(forced_pred)
    st['tdu_pred'] = torch.from_numpy(np.asarray(pred).astype(np.int64))
    seg_label, seg_start, seg_len = run_length(pred)
    sl, ln = torch.from_numpy(seg_label), torch.from_numpy(seg_len)
    st['seg_label'], st['seg_lens'] = sl, ln
    S = len(seg_len)
    # basic.py:615-625 segment mean over ALL H channels
    seg = torch.zeros(S, frame.shape[1]).index_add_(0, sl, frame) / ln[:, None]
    seg = (gru_bidir_fast if fast_gru else gru_bidir)(sd, p + 'seg_update.', seg)
    seg = linear(torch.relu(seg), sd[p + 'seg_combine.weight'], sd[p + 'seg_combine.bias'])
    seg, st['seg_clogit'] = process_feature(seg, C)
    # blocks.py:454-455: centre = int((start+end)/2), end inclusive
    center = torch.from_numpy((seg_start + (seg_start + seg_len - 1)) // 2)
    seg_pos = frame_pos[center] if frame_pos is not None else None
    action, f2a_logit, f2a_attn = x2y_map(sd, p + 'f2a_layer.', seg, action, seg_pos, action_pos)
    action = action_branch(sd, p + 'action_branch.', bc, action, action_pos)
    action, st['action_clogit'] = process_feature(action, C + 1)
    seg_out, a2f_logit, a2f_attn = x2y_map(sd, p + 'a2f_layer.', action, seg, action_pos, seg_pos)
    # blocks.py:439-447: cat[s2f, frame] order matters
    frame = torch.relu(linear(torch.cat([seg_out[sl], frame], -1), sd[p + 'sf_merge.0.weight'], sd[p + 'sf_merge.0.bias']))
    frame = frame_branch(sd, p + 'frame_branch.', frame, bc, False)
    frame, st['frame_clogit'] = process_feature(frame, C)
    st['f2a_attn_logit'], st['a2f_attn_logit'] = f2a_logit, a2f_logit          # (M,S), (S,M)
    st['f2a_attn'] = f2a_attn[:, sl]                                          # blocks.py:481 (M,T)
    st['a2f_attn'] = a2f_attn[sl]                                             # blocks.py:483 (T,M)
    return frame, action


# --------------------------------------------------------------------------------------------
# heads


def feature_projection(sd, p, x):
    """blocks.py:153-175: Linear -> LayerNorm -> ReLU -> (Dropout) -> Linear -> F.normalize."""
    h = linear(x, sd[p + 'projection.0.weight'], sd[p + 'projection.0.bias'])
    h = torch.relu(layer_norm(h, sd[p + 'projection.1.weight'], sd[p + 'projection.1.bias']))
    h = linear(h, sd[p + 'projection.4.weight'], sd[p + 'projection.4.bias'])
    return h / h.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def fuse_eval(action_clogit, a2f_attn, fprob, weight):
    """blocks.py:242-261 (_eval) == blocks.py:854-887 with ``fprob`` = the frame-branch probability."""
    cpred = action_clogit.argmax(1)
    null = action_clogit.shape[-1] - 1
    loc = torch.where(cpred != null)[0]
    if len(loc) == 0:
        return fprob.argmax(1)
    qtk = torch.softmax(action_clogit[:, :-1], dim=1)
    apred = loc[a2f_attn[:, loc].argmax(-1)]
    return ((1 - weight) * qtk[apred] + weight * fprob).argmax(1)


# --------------------------------------------------------------------------------------------
# whole model


def transcript_of(label):
    """basic.py:38-54 (torch_class_label_to_segment_label)[0]: the run-length transcript of a frame label sequence."""
    return torch.unique_consecutive(torch.as_tensor(label).long())


def eval_w_transcript(transcript, a2f_attn, frame_clogit, weight):
    """blocks.py:263-275: frame-branch probability restricted to the transcript classes, softmax over the (already
    normalised) token attention, mixture, argmax over the transcript positions."""
    fprob = torch.softmax(frame_clogit, -1)[:, transcript]
    aprob = torch.softmax(a2f_attn[:, :len(transcript)], -1)
    return transcript[((1 - weight) * aprob + weight * fprob).argmax(1)]


def forward_video(sd, hp, seq, clip=False, forced_preds=None, fast_gru=False, transcript=None):
    """FACT._forward_one_video + eval (blocks.py:56-88,118) or the FACT_CLIP equivalents
    (blocks.py:610-675, 788-887), eval mode.  ``FACT.trans`` True needs ``transcript`` (blocks.py:74-79: the tokens are
    action_embed(transcript) + action_pe, the query position is zero, and vanilla FACT ends in _eval_w_transcript).

    sd: reference-keyed state_dict (fp32 CPU tensors); seq: (T, in_dim).
    Returns a dict: 'blocks' (list of per-block stashes), 'pred' (T,) int64 and, for clip,
    'projected_frame_embeddings' (T,512), 'clip_logit' (T,C).
    """
    C, H = hp['n_classes'], hp['blocks'][0]['hid_dim']
    T = seq.shape[0]
    frame_pos = positional_table(H, max(T, 1)) if hp['fpos'] else None        # basic.py:114-129
    # NB basic.py:93 leaves the table all-zero when FACT.fpos is False; adding zeros == no-op.
    if hp['trans']:
        assert transcript is not None, 'FACT.trans=True needs the transcript'
        A = sd['action_embed.weight'].shape[1]
        action = sd['action_embed.weight'][transcript] + positional_table(A, len(transcript))     # blocks.py:75-78
        action_pos = torch.zeros_like(action)
        frame = seq
    else:
        action_pos = sd['action_query'][:, 0]
        frame, action = seq, torch.zeros_like(action_pos)
    out = {'blocks': []}
    u = 0
    for i, bc in enumerate(hp['blocks']):
        st, p = {}, f'block_list.{i}.'
        if bc['type'] == 'i':
            frame, action = input_block(sd, p, bc, C, frame, action, frame_pos, action_pos, st)
        elif bc['type'] == 'u':
            frame, action = update_block(sd, p, bc, C, frame, action, frame_pos, action_pos, st)
        else:
            fp = None if forced_preds is None else forced_preds[u]
            frame, action = update_block_tdu(sd, p, bc, C, frame, action, frame_pos, action_pos, st, fp, fast_gru)
            u += 1
        st['frame_feature'], st['action_feature'] = frame, action
        out['blocks'].append(st)
    last = out['blocks'][-1]
    if clip and 'text_embeddings' in sd:
        emb = feature_projection(sd, 'frame_projection.', frame[:, :H - C])     # blocks.py:657-660
        out['projected_frame_embeddings'] = emb
        out['clip_logit'] = emb @ sd['text_embeddings'].t() / hp['temp']         # blocks.py:822-823
        fprob = torch.softmax(out['clip_logit'], -1)
    else:
        fprob = torch.softmax(last['frame_clogit'], -1)
    if hp['trans'] and not (clip and 'text_embeddings' in sd):
        out['pred'] = eval_w_transcript(transcript, last['a2f_attn'], last['frame_clogit'], hp['mwt'])
    elif 'a2f_attn' in last:
        out['pred'] = fuse_eval(last['action_clogit'], last['a2f_attn'], fprob, hp['mwt'])
    else:   # a lone input block has no a2f attention; the reference would raise AttributeError
        out['pred'] = fprob.argmax(1)
    return out


def forward(sd, hp, seq_list, clip=False, fast_gru=True):
    """Reference-shaped driver (blocks.py:108-132 / 889-917): list in, list of {'pred': np.int64[T]} out."""
    with torch.no_grad():
        return [{'pred': forward_video(sd, hp, s, clip=clip, fast_gru=fast_gru)['pred'].numpy()} for s in seq_list]
