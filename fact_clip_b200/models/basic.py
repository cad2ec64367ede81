"""Building blocks of the FACT drop-in (counterpart of fact_clip/models/basic.py in the reference).

These modules own the parameters under exactly the reference's names and shapes (SURVEY.md
Appendix B), so ``state_dict`` round-trips with reference checkpoints in both directions, and they
are initialised by the same torch layer constructors in the same order (same values under the same
seed).  They carry NO math of their own: ``forward`` of every class hands a single video to the
factk engine (fact_clip_b200/engine.py -> libfactk.so).  There is no PyTorch fallback.
"""
import math

import torch
import torch.nn as nn


class PositionalEncoding(nn.Module):
    """Sinusoid table buffer 'pe' (max_len, 1, d_model); all zeros when ``empty`` (basic.py:67-129)."""

    def __init__(self, d_model, max_len=5000, empty=False):
        super().__init__()
        self.d_model, self.max_len, self.empty = d_model, max_len, empty
        self._build(max_len)

    def _build(self, n):
        pe = torch.zeros(n, self.d_model)
        if not self.empty:
            pos = torch.arange(0, n, dtype=torch.float).unsqueeze(1)
            div = torch.exp(torch.arange(0, self.d_model, 2).float() * (-math.log(10000.0) / self.d_model))
            pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
        self.register_buffer('pe', pe.unsqueeze(1))

    def forward(self, x):
        if x.size(0) > self.pe.shape[0]:            # table regrowth past max_len (basic.py:125-127)
            dev = self.pe.device
            self._build(x.size(0) + 10)
            self.pe = self.pe.to(dev)
        return self.pe[:x.size(0)]

    def __repr__(self):
        return 'PositionalEncoding(EMPTY)' if self.empty else f'PositionalEncoding(Dim={self.d_model}, MaxLen={self.max_len})'


class DilatedResidualLayer(nn.Module):
    def __init__(self, dilation, nchannels, dropout=0.5, layernorm=True, layernorm_eps=1e-5, ngroup=1):
        super().__init__()
        self.dilation, self.nchannels, self.dropout_rate = dilation, nchannels, dropout
        self.conv_dilated = nn.Conv1d(nchannels, nchannels, 3, padding=dilation, dilation=dilation, groups=ngroup)
        self.conv_1x1 = nn.Conv1d(nchannels, nchannels, 1)
        self.use_layernorm, self.ngroup = layernorm, ngroup
        self.norm = nn.LayerNorm(nchannels, eps=layernorm_eps) if layernorm else None

    def __repr__(self):
        return (f'DilatedResidualLayer(Conv(d={self.dilation},h={self.nchannels}), 1x1(h={self.nchannels}), '
                f'Dropout={self.dropout_rate}, ln={self.use_layernorm})')


class _FrameBranch(nn.Module):
    """Parameter holder for MSTCN / MSTCN2; the whole-model forward drives the engine directly."""

    def forward(self, x, mask=None):
        raise NotImplementedError('sub-module forward is not part of the hot path: call the FACT / FACT_CLIP model '
                                  '(fact_clip_b200.engine.FactEngine.frame_branch runs this branch on device)')


class MSTCN(_FrameBranch):
    def __init__(self, in_dim, hid_dim, out_dim, num_layers, dropout=0.5, dilation_factor=2, ln=True, ngroup=1, in_map=False):
        super().__init__()
        if in_map:
            self.conv_1x1 = nn.Conv1d(in_dim, hid_dim, 1)
        else:
            assert in_dim == hid_dim
        self.layers = nn.ModuleList([DilatedResidualLayer(dilation_factor ** i, hid_dim, dropout, layernorm=ln, ngroup=ngroup)
                                     for i in range(num_layers)])
        self.conv_out = nn.Conv1d(hid_dim, out_dim, 1)
        self.string = (f'MSTCN(h:{in_dim}->{hid_dim}x{num_layers}->{out_dim}, d={dilation_factor}, ng={ngroup}, '
                       f'dropout={dropout}, in_map={in_map})')

    def __repr__(self):
        return self.string


class MSTCN2(_FrameBranch):
    def __init__(self, dim, num_f_maps, out_dim, num_layers, dropout=0.5, dilation_factor=2, ngroup=1, ln=False, in_map=True):
        super().__init__()
        assert not ln, 'MSTCN++ has no layer norm (basic.py:226)'
        self.num_layers, self.ngroup = num_layers, ngroup
        if in_map:
            self.conv_1x1_in = nn.Conv1d(dim, num_f_maps, 1)
        else:
            assert dim == num_f_maps
        d = dilation_factor
        self.conv_dilated_1 = nn.ModuleList(nn.Conv1d(num_f_maps, num_f_maps, 3, padding=d ** (num_layers - 1 - i),
                                                      dilation=d ** (num_layers - 1 - i), groups=ngroup)
                                            for i in range(num_layers))
        self.conv_dilated_2 = nn.ModuleList(nn.Conv1d(num_f_maps, num_f_maps, 3, padding=d ** i, dilation=d ** i, groups=ngroup)
                                            for i in range(num_layers))
        self.conv_fusion = nn.ModuleList(nn.Conv1d(2 * num_f_maps, num_f_maps, 1) for i in range(num_layers))
        self.conv_out = nn.Conv1d(num_f_maps, out_dim, 1)
        self.string = (f'MSTCN2(h:{dim}->{num_f_maps}x{num_layers}->{out_dim}, d={dilation_factor}, ng={ngroup}, '
                       f'dropout={dropout}, in_map={in_map})')

    def __repr__(self):
        return self.string


class X2Y_map(nn.Module):
    """Single-head cross attention between frames/segments and action tokens (basic.py:335-389)."""

    def __init__(self, x_dim, y_dim, y_outdim, head_dim, dropout=0.5, kq_pos=False):
        super().__init__()
        assert kq_pos, 'the model always builds X2Y_map with kq_pos=True (blocks.py:234-238)'
        self.X_K = nn.Linear(x_dim, head_dim)
        self.X_V = nn.Linear(x_dim, head_dim)
        self.Y_Q = nn.Linear(y_dim, head_dim)
        self.Y_W = nn.Linear(y_dim + head_dim, y_outdim)


class SALayer(nn.Module):
    def __init__(self, q_dim, nhead, dim_feedforward=2048, kv_dim=None, dropout=0.1, attn_dropout=0.1):
        super().__init__()
        kv_dim = q_dim if kv_dim is None else kv_dim
        self.multihead_attn = nn.MultiheadAttention(q_dim, nhead, kdim=kv_dim, vdim=kv_dim, dropout=attn_dropout)
        self.linear1 = nn.Linear(q_dim, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, q_dim)
        self.norm1 = nn.LayerNorm(q_dim)
        self.norm2 = nn.LayerNorm(q_dim)
        self.desc = f'SALayer( q({q_dim})xkv({kv_dim})->{q_dim}, head:{nhead}, ffdim:{dim_feedforward}, dropout:{(dropout, attn_dropout)} )'

    def __repr__(self):
        return self.desc


class SCALayer(nn.Module):
    def __init__(self, action_dim, frame_dim, nhead, dim_feedforward=2048, dropout=0.1, attn_dropout=0.1):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(action_dim, nhead, dropout=attn_dropout)
        self.multihead_attn = nn.MultiheadAttention(action_dim, nhead, kdim=frame_dim, vdim=frame_dim, dropout=attn_dropout)
        self.linear1 = nn.Linear(action_dim, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, action_dim)
        self.norm1 = nn.LayerNorm(action_dim)
        self.norm2 = nn.LayerNorm(action_dim)
        self.norm3 = nn.LayerNorm(action_dim)
        self.desc = f'SCALayer( adim:{action_dim}, fdim:{frame_dim}, head:{nhead}, ffdim:{dim_feedforward}, dropout:{(dropout, attn_dropout)} )'

    def __repr__(self):
        return self.desc


def _clones(layer, n):
    import copy
    return nn.ModuleList([copy.deepcopy(layer) for _ in range(n)])


class SCADecoder(nn.Module):
    def __init__(self, in_dim, hid_dim, out_dim, decoder_layer, num_layers, norm=None, in_map=False):
        super().__init__()
        assert not in_map and hid_dim == in_dim
        self.layers = _clones(decoder_layer, num_layers)
        self.out_linear = nn.Linear(hid_dim, out_dim)
        self.num_layers, self.norm = num_layers, norm


class ActionUpdate_GRU(nn.Module):
    """Parameter holder of the GRU action branch for transcript-conditioned models (basic.py:283-308): a bidirectional
    multi-layer nn.GRU over the tokens, LayerNorm, optional output map."""

    def __init__(self, in_dim, hid_dim, out_dim, n_layers, dropout=0.5, layer_norm_eps=1e-5, out_map=False):
        super().__init__()
        self.in_dim, self.hid_dim, self.n_layers = in_dim, hid_dim, n_layers
        self.gru = nn.GRU(in_dim, hid_dim // 2, n_layers, dropout=dropout, bidirectional=True)
        self.layernorm = nn.LayerNorm(hid_dim, eps=layer_norm_eps)
        if out_map:
            self.out_map = nn.Linear(hid_dim, out_dim)
        else:
            assert hid_dim == out_dim
            self.out_map = nn.Identity()


class SADecoder(nn.Module):
    def __init__(self, in_dim, hid_dim, out_dim, decoder_layer, num_layers, norm=None, in_map=False):
        super().__init__()
        assert not in_map and hid_dim == in_dim and norm is None
        self.layers = _clones(decoder_layer, num_layers)
        self.out_linear = nn.Linear(hid_dim, out_dim)
        self.num_layers, self.norm = num_layers, norm
