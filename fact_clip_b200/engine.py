"""Host-side orchestration of the FACT / FACT_CLIP forward over the factk kernels.

One call runs a whole BATCH of videos: frames are packed as [B, slot, C] "rows" tensors (slot = the
longest video rounded up to 128), tokens as [B, M, C], segments reuse the frame slots with a
device-side count per video.  Nothing in here synchronises with the host: the data-dependent
segment count S of the TDU blocks (models/blocks.py:420-426 does a D2H + Python parse) stays on the
device, grids are sized for the worst case and exit early.

Algebraic restructuring used (exact in real arithmetic, fp32 rounding-level differences only):
  * X2Y_map in the f2a direction (models/basic.py:349-389): K/V projections of all T frames are never
    formed.  logit = (Wk^T yq) . (x + pos) + yq.bk, and attn @ (X Wv^T + bv) = (attn @ X) Wv^T + bv.
  * X2Y_map in the a2f direction: logit = (x + pos) . (Wq^T xk) + bq.xk, and
    Y_W [y, attn @ xv] = y Wy^T + attn @ (xv Wa^T).
"""
import math
import os

import torch

from . import ops
from .ops import S


def _round_up(x, m):
    return (x + m - 1) // m * m


class FactEngine:
    def __init__(self, module, hp, clip, mode='bf16'):
        assert mode in ('bf16', 'fp32')
        self.m, self.hp, self.clip, self.mode = module, hp, clip, mode
        self.act = torch.bfloat16 if mode == 'bf16' else torch.float32
        # activation buffers live in one ARENA per batch shape (B, slot); the arenas form an LRU (max_arenas) and an evicted
        # arena takes the CUDA graphs captured over its pointers with it, so a sweep over videos of many different lengths
        # holds a bounded amount of HBM (~25 KB per frame slot of the newest shapes) instead of one arena per shape ever seen
        self._arenas, self._arena, self.max_arenas = {}, None, 4
        self._set_arena(('init',))
        self.use_tc = True
        self.use_fused_tcn = True
        self.use_pair_gemm = True
        self.use_fused_x2y = os.environ.get('FACTK_FUSED_X2Y', '1') != '0'
        self.use_merged_kv = os.environ.get('FACTK_MERGED_KV', '1') != '0'
        self._submits, self._copy_stream, self._slot_free, self._slot_pending = 0, None, [None, None], [None, None]
        self._wcache, self._wsig = {}, None
        # FACTK_FLAT_TOKENS=1 tiles the token rows of all videos as ONE dense matrix (38 instead of 64 tiles at 64 x 75 tokens).
        # Measured: 1.46 instead of 1.41 ms per step for the 73 token GEMMs -- fewer, fuller tiles stream more per CTA and the
        # launches are one wave either way -- so it stays off.
        self.flat_tokens = os.environ.get('FACTK_FLAT_TOKENS', '0') == '1'
        # activation arenas (and the graphs captured over them) are per (batch shape, lane): two batches in flight on two streams
        # -- one filling the SMs that the other's latency-bound GRU chain leaves idle -- use lanes 0 and 1
        self.lane = 0
        # one launch per token-side decoder (half-)layer (csrc/token_layer.cu) instead of 8-11 dependent ones; bf16 mode only
        self.use_fused_tokens = os.environ.get('FACTK_FUSED_TOKENS', '1') != '0'
        self.use_graph = True        # replay the whole batched forward as ONE CUDA graph (no per-kernel host launch cost)
        # Epic verb/noun model (blocks_SepVerbNoun.py): two class heads, action table (verb id, noun id) per action
        self.vn = hp.get('vn')
        # token state of a call: query-token models always use these; FACT.trans models set them per video in run()
        self.ntok, self.action_init, self.transcript = hp['ntoken'], None, None
        self.last_launches = 0

    # ------------------------------------------------------------------ memory / weights
    def _set_arena(self, ctx):
        """Make the arena of batch shape ``ctx`` current (most recently used); evict the least recently used beyond
        ``max_arenas`` together with every graph captured over it."""
        a = self._arenas.pop(ctx, None)
        if a is None:
            a = dict(bufs={}, zbufs={}, len_sig=None, graphs={}, warmed=False)
        self._arenas[ctx] = a                       # dicts keep insertion order: last = most recent
        while len(self._arenas) > self.max_arenas:
            self._arenas.pop(next(iter(self._arenas)))
        self._arena = a
        self._bufs, self._zbufs = a['bufs'], a['zbufs']

    @property
    def _len_sig(self):
        return self._arena['len_sig']

    @_len_sig.setter
    def _len_sig(self, v):
        self._arena['len_sig'] = v

    @property
    def _graphs(self):
        return self._arena['graphs']

    def buf(self, name, shape, dtype=torch.float32):
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            # zero-initialised ONCE: rows in [len, slot) are never written, and the tensor-core attention kernels multiply them by
            # exact-zero probabilities -- stale finite values are harmless there, uninitialised NaN bit patterns would not be
            t = torch.zeros(shape, dtype=dtype, device=self.dev)
            if not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream().synchronize()  # (the fill must not race a copy-stream write into the new buffer)
            self._bufs[key] = t
        return t

    def zbuf(self, name, shape, dtype):
        """Buffer that feeds convolution taps: rows in [len, slot) must read as zero for the TMA path, so it is
        zero-initialised and re-zeroed whenever the batch's lengths change (no kernel ever writes those rows)."""
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.dev)
            self._bufs[key] = t
            self._zbufs[key] = t
        return t

    def wbf(self, W):
        """bf16 copy of an fp32 weight view (cached until the master parameters change)."""
        key = ('bf16', W.data_ptr(), tuple(W.shape), tuple(W.stride()))
        return self.derived(key, lambda: W.to(torch.bfloat16))

    def x2y_tc(self, rows, H):
        return self.mode == 'bf16' and self.use_tc and rows.dtype == torch.bfloat16 and H % 64 == 0

    def with_pos(self, rows, rlen, pos_idx, name):
        """add_positional_encoding (basic.py:313-320) on frame / segment rows for the tensor-core path: ``rows + frame_pos[t or
        centre]`` materialised ONCE (one elementwise pass, bf16) and shared by every GEMM that reads positioned rows -- the six
        key projections of the SCA decoder, the f2a and a2f logits of an update block -- so that FACT.fpos models (Epic) stay
        on the tcgen05 kernels instead of the CUDA-core GEMM with a per-element positional term."""
        if self.frame_pos is None:
            return rows
        y = self.zbuf(name, tuple(rows.shape), rows.dtype)
        ops.ew(ops.EW_ADDTAB, rows, y, rows.shape[-1], r=self.frame_pos, len=rlen, ridx=pos_idx)
        return y

    def lin(self, x, W, N, out, pos=None, **kw):
        """Token-side Linear with an optional query-position on the input.  On the tensor-core path (tf32) the
        position term moves to a cached table: (x + pos) W^T = x W^T + (pos W[:, :d]^T)."""
        K = W.shape[-1]
        if (self.mode == 'bf16' and self.use_tc and x.dtype == torch.float32 and W.dim() == 2 and K % 32 == 0
                and x.shape[0] == out.shape[0]):
            pre = None
            if pos is not None:
                assert kw.get('alpha', 1.0) == 1.0
                d = pos.shape[-1]
                pre = self.derived(('posW', W.data_ptr(), tuple(W.shape), tuple(W.stride())), lambda: pos @ W[:, :d].t())
            if self.flat_tokens and x.dim() == 3 and x.shape[0] > 1:
                # the token rows of all videos form one dense matrix: tile it as ONE problem (B * M rows -> ceil(B M / 128)
                # tiles) instead of one 128-row tile per video with M valid rows
                flat = self._flatten_rows(x, out, kw.get('res'))
                if flat is not None:
                    xf, of, rf = flat
                    kw2 = dict(kw)
                    if rf is not None:
                        kw2['res'] = rf
                    Bv, M = x.shape[0], x.shape[1]
                    pidx = None
                    if pre is not None:
                        pidx = self.derived(('tokidx', Bv, M), lambda: (torch.arange(Bv * M, device=self.dev, dtype=torch.int32) % M)[None])
                    return ops.gemm([S(xf, W)], N, of, tc=True, pre=pre, pre_idx=pidx, tag='tok_tc', **kw2)
            return ops.gemm([S(x, W)], N, out, tc=True, pre=pre, tag='tok_tc', **kw)
        return ops.gemm([S(x, W, pos=pos)], N, out, **kw)

    @staticmethod
    def _flatten_rows(x, out, res):
        """[B, M, *] views whose rows are equally spaced across the batch -> [1, B*M, *] views (None if any is not)."""
        def flat(t):
            if t is None:
                return None
            Bv, M = t.shape[0], t.shape[1]
            if t.stride(-1) != 1 or t.stride(0) != M * t.stride(1):
                return False
            return torch.as_strided(t, (1, Bv * M, t.shape[2]), (Bv * M * t.stride(1), t.stride(1), 1), t.storage_offset())
        xf, of, rf = flat(x), flat(out), flat(res)
        if xf is False or of is False or rf is False:
            return None
        return xf, of, rf

    def mm(self, srcs, N, out, tf32=False, **kw):
        """GEMM dispatch: tcgen05 kernel when the operands qualify (bf16 mode), CUDA-core kernel otherwise."""
        if self.mode == 'bf16' and self.use_tc:
            s0 = srcs[0]
            if (self.use_pair_gemm and len(srcs) == 1 and s0['gather'] is None and s0['pos'] is None and s0['off'] == 0 and s0['K'] is None
                    and s0['W'].dim() == 2 and s0['A'].dtype == torch.bfloat16 and out.dtype == torch.bfloat16
                    and kw.get('res') is None and kw.get('alpha', 1.0) == 1.0 and (kw.get('bias') is None or kw['bias'].dim() == 1)
                    and s0['W'].shape[1] % 64 == 0 and N % 128 == 0):
                Wb = self.wbf(s0['W'])
                if ops.gemm_pair_ok(s0['A'], Wb, N, out, kw.get('pre')):
                    return ops.gemm_pair(s0['A'], Wb, N, out, len=kw.get('len'), bias=kw.get('bias'), relu=kw.get('relu', False),
                                         pre=kw.get('pre'), pre_idx=kw.get('pre_idx'), tag=kw.get('tag'))
            dts = {s['A'].dtype for s in srcs}
            plain = all(s['gather'] is None and s['pos'] is None and s['W'].dim() == 2 for s in srcs)
            if plain and dts == {torch.bfloat16} and all((s['K'] or s['W'].shape[-1]) % 64 == 0 for s in srcs):
                for s in srcs:
                    s['W'] = self.wbf(s['W'])
                return ops.gemm(srcs, N, out, tc=True, **kw)
            if plain and tf32 and dts == {torch.float32} and all((s['K'] or s['W'].shape[-1]) % 32 == 0 for s in srcs):
                return ops.gemm(srcs, N, out, tc=True, **kw)
        return ops.gemm(srcs, N, out, **kw)

    def _refresh_weights(self):
        params = dict(self.m.named_parameters())
        params.update(dict(self.m.named_buffers()))
        sig = tuple((k, v.data_ptr(), v._version) for k, v in params.items())
        if sig != self._wsig:
            self._wsig, self._wcache, self._p = sig, {}, params
            for a in self._arenas.values():       # captured graphs hold pointers of derived weights
                a['graphs'] = {}
        dev = next(self.m.parameters()).device
        if getattr(self, 'dev', dev) != dev:      # the module moved to another GPU: every cached buffer lives on the old one
            self._arenas = {}
            self._set_arena(('init',))
            self._copy_stream, self._slot_free, self._slot_pending = None, [None, None], [None, None]
        self.dev = dev

    def p(self, name):
        t = self._p[name]
        assert t.dtype == torch.float32 and t.is_cuda, f'{name}: fp32 CUDA master parameter expected'
        return t.detach()

    def derived(self, key, fn):
        t = self._wcache.get(key)
        if t is None:
            with torch.no_grad():
                t = fn().contiguous()
            self._wcache[key] = t
        return t

    def taps(self, name, groups=1):
        """Conv1d weight (Cout, Cin/groups, k) -> [k][Cout][Cin].  A grouped convolution (f_ngp > 1, basic.py:139) becomes
        its block-diagonal dense matrix so every GEMM path (fused layer, folded MSTCN++ taps) serves it unchanged."""
        if groups == 1:
            return self.derived(('taps', name), lambda: self.p(name).permute(2, 0, 1))

        def dense():
            w = self.p(name)
            co, ci = w.shape[0] // groups, w.shape[1]
            full = w.new_zeros(w.shape[0], ci * groups, w.shape[2])
            for j in range(groups):
                full[j * co:(j + 1) * co, j * ci:(j + 1) * ci] = w[j * co:(j + 1) * co]
            return full.permute(2, 0, 1)
        return self.derived(('taps', name, groups), dense)

    def tr(self, name):        # W^T
        return self.derived(('tr', name), lambda: self.p(name).t())

    def cat(self, *names):
        return self.derived(('cat',) + names, lambda: torch.cat([self.p(n) for n in names], 0))

    # ------------------------------------------------------------------ class heads
    def ncls(self, tokens=False):
        """Width of the class-logit tail of a feature row: C (+1 null class for tokens), or both verb/noun heads."""
        if self.vn is None:
            return self.hp['n_classes'] + (1 if tokens else 0)
        n1, n2 = self.hp['n_classes']
        return n1 + n2 + (2 if tokens else 0)

    def vn_table(self):
        mk = lambda k: torch.tensor(self.vn[k], dtype=torch.int32, device=self.dev)
        return self.derived(('vn_vids',), lambda: mk(0)), self.derived(('vn_nids',), lambda: mk(1))

    def splice(self, x, clogit, len=None, tokens=False, pred=None):
        """Block.process_feature in place on the class tail of x (blocks.py:195-202; blocks_SepVerbNoun.py:229-234 for the
        two-head model) + the segmentation argmax for the next TDU block when ``pred`` is given."""
        if self.vn is None:
            return ops.softmax_splice(x, self.ncls(tokens), clogit, pred, len=len)
        n1, n2 = self.hp['n_classes']
        vids, nids = self.vn_table()
        k = 1 if tokens else 0
        ops.vn_splice(x, n1 + k, n2 + k, clogit, vids if pred is not None else None, nids if pred is not None else None,
                      pred, len=len)

    # ------------------------------------------------------------------ frame branch
    def frame_branch(self, pfx, bc, x, in_map, tag):
        """MSTCN / MSTCN2 (models/basic.py:200-220, 263-281) + process_feature (models/blocks.py:195-202).
        x: [B, slot, Din].  Returns (frame_feature [B,slot,H] act dtype, frame_clogit fp32, pred int32)."""
        B, slot, F, H, Lr, C = self.B, self.slot, bc['f_dim'], bc['hid_dim'], bc['f_layers'], self.ncls()
        ln = self.len
        fa, fb = self.zbuf('f_a', (B, slot, F), self.act), self.zbuf('f_b', (B, slot, F), self.act)
        m2, ng = bc['f'] == 'm2', bc['f_ngp']
        other = lambda t: fb if t is fa else fa
        if in_map:
            w = pfx + ('conv_1x1_in' if m2 else 'conv_1x1')
            self.mm([S(x, self.taps(w + '.weight')[0])], F, fa, tf32=True, len=ln, bias=self.p(w + '.bias'), tag='in_proj')
            cur, nxt = fa, fb
        else:
            cur, nxt = x, fa
        fused = (not m2 and self.mode == 'bf16' and self.use_tc and self.use_fused_tcn and ops.tcn_layer_supported(F)
                 and cur.dtype == torch.bfloat16 and cur.is_contiguous())
        for i in range(Lr):
            if fused:
                # conv3 + ReLU + 1x1 + residual in one tcgen05 kernel; the ReLU tile never leaves the SM (tcn_fused.cu)
                q = f'{pfx}layers.{i}.'
                ops.tcn_layer(cur, nxt, self.wbf(self.taps(q + 'conv_dilated.weight', ng)), self.p(q + 'conv_dilated.bias'),
                              self.wbf(self.taps(q + 'conv_1x1.weight')[0]), self.p(q + 'conv_1x1.bias'), 2 ** i, len=ln)
            elif not m2:
                q = f'{pfx}layers.{i}.'
                w3, d = self.taps(q + 'conv_dilated.weight', ng), 2 ** i
                tmp = self.buf('f_tmp', (B, slot, F), self.act)
                self.mm([S(cur, w3[k], off=(k - 1) * d) for k in range(3)], F, tmp, len=ln,
                        bias=self.p(q + 'conv_dilated.bias'), relu=True, tag='tcn_conv3')
                self.mm([S(tmp, self.taps(q + 'conv_1x1.weight')[0])], F, nxt, len=ln,
                        bias=self.p(q + 'conv_1x1.bias'), res=cur, tag='tcn_1x1')
            else:
                # MSTCN2 layer (basic.py:271-277): conv_fusion(cat[conv_d1(x), conv_d2(x)]) has no non-linearity inside, so it
                # folds into ONE 5-tap convolution with weights Wf1 W1_k / Wf2 W2_k (centre taps merged): 10 F^2 instead of
                # 16 F^2 FLOP per frame, one launch instead of three, no 2F-wide intermediate.
                offs = self._m2_offsets(i, Lr)
                Wt = self.derived(('m2fold_w', pfx, i), lambda: self._m2_fold(pfx, i, Lr, ng)[0])
                bt = self.derived(('m2fold_b', pfx, i), lambda: self._m2_fold(pfx, i, Lr, ng)[1])
                self.mm([S(cur, Wt[j], off=o) for j, o in enumerate(offs)], F, nxt, len=ln, bias=bt, relu=True, res=cur, tag='tcn_m2')
            if bc['f_ln'] and not m2:      # DilatedResidualLayer.norm (basic.py:166-169): LayerNorm over the channels, in place
                q = f'{pfx}layers.{i}.'
                ops.layernorm(nxt, self.p(q + 'norm.weight'), self.p(q + 'norm.bias'), nxt, len=ln)
            cur, nxt = nxt, other(nxt)
        out = self.buf('frame_' + tag, (B, slot, H), self.act)
        self.mm([S(cur, self.taps(pfx + 'conv_out.weight')[0])], H, out, len=ln, bias=self.p(pfx + 'conv_out.bias'), tag='conv_out')
        clogit = self.buf('fclogit_' + tag, (B, slot, C))
        pred = self.buf('fpred_' + tag, (B, slot), torch.int32)
        self.splice(out, clogit, len=ln, pred=pred)
        return out, clogit, pred

    @staticmethod
    def _m2_offsets(i, Lr):
        d1, d2 = 2 ** (Lr - 1 - i), 2 ** i
        return sorted({-d1, 0, d1, -d2, d2})

    def _m2_fold(self, pfx, i, Lr, groups=1):
        """Folded taps [n_offsets, F, F] (in _m2_offsets order) and bias [F] of MSTCN2 layer i."""
        w1 = self.taps(f'{pfx}conv_dilated_1.{i}.weight', groups).permute(1, 2, 0)                              # (F, F, 3)
        w2 = self.taps(f'{pfx}conv_dilated_2.{i}.weight', groups).permute(1, 2, 0)
        b1, b2 = self.p(f'{pfx}conv_dilated_1.{i}.bias'), self.p(f'{pfx}conv_dilated_2.{i}.bias')
        wf, bf = self.p(f'{pfx}conv_fusion.{i}.weight')[:, :, 0], self.p(f'{pfx}conv_fusion.{i}.bias')       # (F, 2F)
        F = w1.shape[0]
        A1, A2 = wf[:, :F].double(), wf[:, F:].double()
        d1, d2 = 2 ** (Lr - 1 - i), 2 ** i
        acc = {}
        for k in range(3):
            acc[(k - 1) * d1] = acc.get((k - 1) * d1, 0) + A1 @ w1[:, :, k].double()
            acc[(k - 1) * d2] = acc.get((k - 1) * d2, 0) + A2 @ w2[:, :, k].double()
        W = torch.stack([acc[o] for o in self._m2_offsets(i, Lr)]).float().contiguous()
        b = (A1 @ b1.double() + A2 @ b2.double() + bf.double()).float().contiguous()
        return W, b

    # ------------------------------------------------------------------ token side
    def _tok_fused(self, x, nhead, ff):
        B, M, A = x.shape
        return (self.mode == 'bf16' and self.use_tc and self.use_fused_tokens and x.is_contiguous() and x.dtype == torch.float32
                and ops.token_layer_ok(M, A, nhead, ff))

    def _tokw(self, W):
        """Weight in the fused token kernel's packed bf16 fragment order (cached like every derived weight)."""
        return self.derived(('tokw', W.data_ptr(), tuple(W.shape), tuple(W.stride())), lambda: ops.pack_token_weight(W))

    def _posw(self, W, pos):
        """(x + pos) W^T = x W^T + pos W[:, :d]^T: the position term as a table (shared with lin())."""
        if pos is None:
            return None
        d = pos.shape[-1]
        return self.derived(('posW', W.data_ptr(), tuple(W.shape), tuple(W.stride())), lambda: pos @ W[:, :d].t())

    def _ffn_args(self, q, n_a, n_b):
        w1, w2 = self.p(q + 'linear1.weight'), self.p(q + 'linear2.weight')
        return (self._tokw(w1), self.p(q + 'linear1.bias'), self._tokw(w2), self.p(q + 'linear2.bias'), self.p(q + n_a), self.p(q + n_b),
                w1.shape[0])

    def _mha_self(self, pfx, x, pos, nhead, tag):
        """q = k = x + pos, v = x through a packed in_proj (basic.py:437,500); returns attn output before out_proj."""
        B, M, A = x.shape
        W, bias = self.p(pfx + 'in_proj_weight'), self.p(pfx + 'in_proj_bias')
        qkv = self.buf('tok_qkv', (B, M, 3 * A))
        self.lin(x, W[:2 * A], 2 * A, qkv[:, :, :2 * A], pos=pos, bias=bias[:2 * A])
        self.lin(x, W[2 * A:], A, qkv[:, :, 2 * A:], bias=bias[2 * A:])
        # (one merged q|k|v launch was measured slower: 192 tiles = two waves on 148 SMs, 45 us against 19 + 14 us)
        o = self.buf('tok_o', (B, M, A))
        ops.mha_tokens(qkv[:, :, :A], qkv[:, :, A:2 * A], qkv[:, :, 2 * A:], o, nhead, tf32=(self.mode == 'bf16' and self.use_tc))
        return o

    def _ffn_ln(self, q, x, n_a, n_b, tag):
        B, M, A = x.shape
        ff = self.p(q + 'linear1.weight').shape[0]
        h = self.buf('tok_ff', (B, M, ff))
        self.lin(x, self.p(q + 'linear1.weight'), ff, h, bias=self.p(q + 'linear1.bias'), relu=True)
        t = self.buf('tok_t', (B, M, A))
        self.lin(h, self.p(q + 'linear2.weight'), A, t, bias=self.p(q + 'linear2.bias'), res=x)
        ops.layernorm(t, self.p(q + n_a), self.p(q + n_b), x)

    def sca_decoder(self, pfx, bc, frame, tag, rlen=None, pos_idx=None):
        """SCADecoder over SCALayer (models/basic.py:542-557, 494-523): tokens attend the memory rows ``frame`` (frames; or
        segments with ``rlen`` rows per video and positions frame_pos[pos_idx], blocks_SepVerbNoun.py:380-382). -> [B,M,H] fp32."""
        B, slot, M, A, H, nh = self.B, self.slot, self.ntok, bc['a_dim'], bc['hid_dim'], bc['a_nhead']
        rlen = self.len if rlen is None else rlen
        qpos = self.qpos()
        tgt = self.buf('tok_x', (B, M, A))
        if self.action_init is None:
            tgt.zero_()
        else:
            tgt.copy_(self.action_init)
        t = self.buf('tok_t', (B, M, A))
        ws = self.buf('attn_ws', (max(ops.attn_rows_ws(B, slot, M, nh, A // nh), 1),))
        fpos = self.frame_pos
        ff = self.p(f'{pfx}layers.0.linear1.weight').shape[0]
        fused = self._tok_fused(tgt, nh, ff)

        def xattn_w(i):
            c = f'{pfx}layers.{i}.multihead_attn.'
            if (c + 'in_proj_weight') in self._p:
                W = self.p(c + 'in_proj_weight')
                return W[:A], W[A:2 * A], W[2 * A:], self.p(c + 'in_proj_bias')
            return self.p(c + 'q_proj_weight'), self.p(c + 'k_proj_weight'), self.p(c + 'v_proj_weight'), self.p(c + 'in_proj_bias')

        nl = bc['a_layers']
        kv_all = None
        if fpos is None and self.mode == 'bf16' and self.use_tc and frame.dtype == torch.bfloat16 and nl > 1 and self.use_merged_kv:
            # every layer projects the SAME memory rows: one GEMM with the key / value weights of all layers stacked (N = 2 A layers)
            # reads the rows once instead of once per layer; layer i attends columns [2 A i, 2 A (i + 1)) of the result
            wkv_all = self.derived(('wkv_all', pfx), lambda: torch.cat([torch.cat(xattn_w(i)[1:3], 0) for i in range(nl)], 0))
            bkv_all = self.derived(('bkv_all', pfx), lambda: torch.cat([xattn_w(i)[3][A:] for i in range(nl)], 0))
            kv_all = self.zbuf('sca_kv_all', (B, slot, 2 * A * nl), self.act)
            self.mm([S(frame, wkv_all)], 2 * A * nl, kv_all, len=rlen, bias=bkv_all, tag='sca_kv')
        # zero-initialised: rows >= len are never written and the tensor-core attention multiplies them by exact zeros
        kv = self.zbuf('sca_kv', (B, slot, 2 * A), self.act) if kv_all is None else None
        for i in range(nl):
            q = f'{pfx}layers.{i}.'
            c = q + 'multihead_attn.'
            wq, wk, wv, cb = xattn_w(i)
            if kv_all is not None:
                kv = kv_all[:, :, 2 * A * i:2 * A * (i + 1)]
            cq = self.buf('tok_cq', (B, M, A))
            if fused:
                # self attention + out_proj + norm1 + the cross attention's query projection: one launch
                Wi = self.p(q + 'self_attn.in_proj_weight')
                ops.token_layer(tgt, nh, self._tokw(self.p(q + 'self_attn.out_proj.weight')), self.p(q + 'self_attn.out_proj.bias'),
                                self.p(q + 'norm1.weight'), self.p(q + 'norm1.bias'), w_in=self._tokw(Wi),
                                b_in=self.p(q + 'self_attn.in_proj_bias'), pre_qk=self._posw(Wi[:2 * A], qpos),
                                w_q=self._tokw(wq), b_q=cb[:A], pre_q=self._posw(wq, qpos), cq_out=cq)
            else:
                o = self._mha_self(q + 'self_attn.', tgt, qpos, nh, tag)
                self.lin(o, self.p(q + 'self_attn.out_proj.weight'), A, t, bias=self.p(q + 'self_attn.out_proj.bias'), res=tgt)
                ops.layernorm(t, self.p(q + 'norm1.weight'), self.p(q + 'norm1.bias'), tgt)
                # cross attention: q from tokens, k from frames (+pos), v from frames
                self.lin(tgt, wq, A, cq, pos=qpos, bias=cb[:A])
            if kv_all is not None:
                pass
            elif fpos is None:
                wkv = self.derived(('wkv', c), lambda: torch.cat([wk, wv], 0))
                self.mm([S(frame, wkv)], 2 * A, kv, len=rlen, bias=cb[A:], tag='sca_kv')
            elif self.mode == 'bf16' and self.use_tc and frame.dtype == torch.bfloat16:
                if i == 0:
                    frame_p = self.with_pos(frame, rlen, pos_idx, 'sca_rows_pos')
                self.mm([S(frame_p, wk)], A, kv[:, :, :A], len=rlen, bias=cb[A:2 * A], tag='sca_kv')
                self.mm([S(frame, wv)], A, kv[:, :, A:], len=rlen, bias=cb[2 * A:], tag='sca_kv')
            else:
                ops.gemm([S(frame, wk, pos=fpos, pos_idx=pos_idx)], A, kv[:, :, :A], len=rlen, bias=cb[A:2 * A])
                self.mm([S(frame, wv)], A, kv[:, :, A:], len=rlen, bias=cb[2 * A:])
            o = self.buf('tok_o', (B, M, A))
            ops.attn_rows(cq, kv[:, :, :A], kv[:, :, A:], o, nh, ws, len=rlen)
            if fused:
                # out_proj + norm2 + FFN + norm3: one launch
                ops.token_layer(tgt, nh, self._tokw(self.p(c + 'out_proj.weight')), self.p(c + 'out_proj.bias'),
                                self.p(q + 'norm2.weight'), self.p(q + 'norm2.bias'), o_in=o,
                                ffn=self._ffn_args(q, 'norm3.weight', 'norm3.bias'))
                continue
            self.lin(o, self.p(c + 'out_proj.weight'), A, t, bias=self.p(c + 'out_proj.bias'), res=tgt)
            ops.layernorm(t, self.p(q + 'norm2.weight'), self.p(q + 'norm2.bias'), tgt)
            self._ffn_ln(q, tgt, 'norm3.weight', 'norm3.bias', tag)
        ops.layernorm(tgt, self.p(pfx + 'norm.weight'), self.p(pfx + 'norm.bias'), t)
        out = self.buf('action_' + tag, (B, M, H))
        self.lin(t, self.p(pfx + 'out_linear.weight'), H, out, bias=self.p(pfx + 'out_linear.bias'))
        return out

    def sa_decoder(self, pfx, bc, x, tag):
        """SADecoder over SALayer (models/basic.py:578-593, 429-452). x: [B,M,A] fp32 -> [B,M,H] fp32."""
        B, M, A, H, nh = self.B, self.ntok, bc['a_dim'], bc['hid_dim'], bc['a_nhead']
        qpos = self.qpos()
        t = self.buf('tok_t', (B, M, A))
        fused = self._tok_fused(x, nh, self.p(f'{pfx}layers.0.linear1.weight').shape[0])
        for i in range(bc['a_layers']):
            q = f'{pfx}layers.{i}.'
            if fused:                  # the whole SALayer in one launch
                Wi = self.p(q + 'multihead_attn.in_proj_weight')
                ops.token_layer(x, nh, self._tokw(self.p(q + 'multihead_attn.out_proj.weight')), self.p(q + 'multihead_attn.out_proj.bias'),
                                self.p(q + 'norm1.weight'), self.p(q + 'norm1.bias'), w_in=self._tokw(Wi),
                                b_in=self.p(q + 'multihead_attn.in_proj_bias'), pre_qk=self._posw(Wi[:2 * A], qpos),
                                ffn=self._ffn_args(q, 'norm2.weight', 'norm2.bias'))
                continue
            o = self._mha_self(q + 'multihead_attn.', x, qpos, nh, tag)
            self.lin(o, self.p(q + 'multihead_attn.out_proj.weight'), A, t,
                     bias=self.p(q + 'multihead_attn.out_proj.bias'), res=x)
            ops.layernorm(t, self.p(q + 'norm1.weight'), self.p(q + 'norm1.bias'), x)
            self._ffn_ln(q, x, 'norm2.weight', 'norm2.bias', tag)
        out = self.buf('action_' + tag, (B, M, H))
        self.lin(x, self.p(pfx + 'out_linear.weight'), H, out, bias=self.p(pfx + 'out_linear.bias'))
        return out

    def gru_decoder(self, pfx, bc, x, tag):
        """ActionUpdate_GRU (models/basic.py:283-308; FACT.trans models): stacked bidirectional GRU over the tokens,
        LayerNorm, optional out_map.  x: [B,M,A] fp32 -> [B,M,H] fp32.  The recurrence is the fp32 cluster kernel the
        segment GRU uses in fp32 mode (a transcript is tens of tokens: latency, not throughput)."""
        B, M, A, H = self.B, self.ntok, bc['a_dim'], bc['hid_dim']
        Hh, g = A // 2, pfx + 'gru.'
        ntok = self.buf('agru_n', (B,), torch.int32)
        ntok.fill_(M)
        for l in range(bc['a_layers']):
            gi = self.buf('agru_gi', (B, M, 6 * Hh))
            ops.gemm([S(x, self.cat(f'{g}weight_ih_l{l}', f'{g}weight_ih_l{l}_reverse'))], 6 * Hh, gi,
                     bias=self.cat(f'{g}bias_ih_l{l}', f'{g}bias_ih_l{l}_reverse'))
            y = self.buf(f'agru_y{l % 2}', (B, M, A))
            ops.gru_bidir(gi, self.p(f'{g}weight_hh_l{l}'), self.p(f'{g}bias_hh_l{l}'), self.p(f'{g}weight_hh_l{l}_reverse'),
                          self.p(f'{g}bias_hh_l{l}_reverse'), y, ntok, relu=False)
            x = y
        out = self.buf('action_' + tag, (B, M, H))
        if bc['a'] == 'gru_om':
            t = self.buf('agru_ln', (B, M, A))
            ops.layernorm(x, self.p(pfx + 'layernorm.weight'), self.p(pfx + 'layernorm.bias'), t)
            self.lin(t, self.p(pfx + 'out_map.weight'), H, out, bias=self.p(pfx + 'out_map.bias'))
        else:
            ops.layernorm(x, self.p(pfx + 'layernorm.weight'), self.p(pfx + 'layernorm.bias'), out)
        return out

    def action_branch(self, pfx, bc, x, tag, frame=None, rlen=None, pos_idx=None):
        """Block.create_abranch dispatch (models/blocks.py:215-232)."""
        if bc['a'] in ('gru', 'gru_om'):
            return self.gru_decoder(pfx, bc, self.action_init if frame is not None else x, tag)
        if frame is not None:
            return self.sca_decoder(pfx, bc, frame, tag, rlen=rlen, pos_idx=pos_idx)
        return self.sa_decoder(pfx, bc, x, tag)

    def token_splice(self, action, tag):
        clogit = self.buf('aclogit_' + tag, (self.B, self.ntok, self.ncls(tokens=True)))
        self.splice(action, clogit, tokens=True)
        return clogit

    # ------------------------------------------------------------------ cross attention
    def f2a(self, pfx, bc, rows, rlen, pos_idx, action, tag, want_attn):
        """X2Y_map with X = rows (frames or segments), Y = tokens (basic.py:349-389). -> tokens [B,M,A] fp32,
        attn_logit [B,slot,Mp] (row = x, col = token), attn (or None)."""
        B, slot, M, A, H = self.B, self.slot, self.ntok, bc['a_dim'], bc['hid_dim']
        Mp = _round_up(M, 4)
        qpos = self.qpos()
        alpha = 1.0 / math.sqrt(H)
        tc = self.x2y_tc(rows, H)
        qt = self.buf('x2y_qt16' if tc else 'x2y_qt', (B, M, H), torch.bfloat16 if tc else torch.float32)   # alpha * Wk^T yq
        cb = self.buf('x2y_c', (B, M, 1))                                   # alpha * yq . bk
        fold = self.mode == 'bf16' and self.use_tc
        fused = (tc and fold and not want_attn and self.frame_pos is None and self.use_fused_x2y and rows.is_contiguous()
                 and ops.f2a_fused_ok(M, H, slot))
        if fold:
            # yq = Y_Q(action + pos) only feeds linear maps: fold Y_Q into them (one GEMM instead of two, no yq buffer)
            Wf, bf_, u, c0 = self._x2y_fold(pfx, 'Y_Q', 'X_K', alpha)
            self.lin(action, Wf, H, qt, pos=qpos, bias=bf_)
            if not fused:       # (the per-token logit bias is constant along the rows: the fused softmax never needs it)
                self.lin(action, u, 1, cb, pos=qpos, bias=c0)
        else:
            yq = self.buf('x2y_tokH', (B, M, H))
            self.lin(action, self.p(pfx + 'Y_Q.weight'), H, yq, pos=qpos, bias=self.p(pfx + 'Y_Q.bias'))
            self.lin(yq, self.tr(pfx + 'X_K.weight'), H, qt, alpha=alpha)
            self.lin(yq, self.p(pfx + 'X_K.bias')[None, :], 1, cb, alpha=alpha)
        xbar = self.buf('x2y_xbar', (B, M, H))
        if fused:
            # ONE tcgen05 kernel (f2a_fused.cu): S = qt rows^T in TMEM, softmax over the rows on TMEM lanes, xbar += P rows with the
            # same TMA tile as the second operand -- the rows are read once and the fp32 logits never exist (their per-token
            # bias cb is constant along the rows and cancels).  The logits / attention only leave the SM on the unfused path
            # below, which serves the loss and keep_attn (self.keep).
            logit = attn = None
            ws = self.buf('f2a_ws', (ops.f2a_fused_ws(B, slot, M, H),))
            ops.f2a_fused(rows, qt, xbar, M, ws, len=rlen)
        else:
            logit = self.buf('f2a_logit_' + tag, (B, slot, Mp))
            if tc:
                ops.gemm([S(self.with_pos(rows, rlen, pos_idx, 'x2y_rows_pos'), qt)], M, logit, len=rlen, bias=cb[:, :, 0], tc=True, tag='x2y_rows')
            else:
                ops.gemm([S(rows, qt, pos=self.frame_pos, pos_idx=pos_idx)], M, logit, len=rlen, bias=cb[:, :, 0], tag='x2y_rows')
            attn = self.buf('f2a_attn_' + tag, (B, slot, Mp)) if want_attn else None
            ws = self.buf('col_ws', (ops.col_softmax_ws(B, slot, M, H),))
            ops.col_softmax_apply(logit, rows, xbar, M, ws, attn=attn, len=rlen, E=H)
        W = self.p(pfx + 'Y_W.weight')
        out = self.buf('tok_x', (B, M, A))
        if fold:     # Y_W(cat[action, X_V(xbar)]) = action W1^T + xbar (W2 Wv)^T + (b + W2 bv)
            W2v, b2 = self._yw_fold(pfx)
            self.mm([S(action, W[:, :H]), S(xbar, W2v)], A, out, tf32=True, bias=b2)
        else:
            feat = self.buf('x2y_tokH', (B, M, H))
            self.lin(xbar, self.p(pfx + 'X_V.weight'), H, feat, bias=self.p(pfx + 'X_V.bias'))
            self.mm([S(action, W[:, :H]), S(feat, W[:, H:])], A, out, tf32=True, bias=self.p(pfx + 'Y_W.bias'))
        return out, logit, attn

    def _x2y_fold(self, pfx, first, second, alpha):
        """Two chained Linears with nothing in between, q = first(x) then alpha * q W2 (the transposed use of the second
        weight in the logit, basic.py:367-372) and alpha * q . b2:  W' = alpha W2^T W1, b' = alpha W2^T b1,
        u = alpha W1^T b2, c0 = alpha b1 . b2.  Products formed once in fp64."""
        def make():
            W1, b1 = self.p(f'{pfx}{first}.weight').double(), self.p(f'{pfx}{first}.bias').double()
            W2, b2 = self.p(f'{pfx}{second}.weight').double(), self.p(f'{pfx}{second}.bias').double()
            return [(alpha * W2.t() @ W1).float().contiguous(), (alpha * W2.t() @ b1).float().contiguous(),
                    (alpha * W1.t() @ b2).float()[None, :].contiguous(), (alpha * b1 @ b2).float().reshape(1).contiguous()]
        return [self.derived(('x2yfold', pfx, first, i), lambda i=i: make()[i]) for i in range(4)]

    def _yw_fold(self, pfx):
        def make():
            W, b = self.p(pfx + 'Y_W.weight').double(), self.p(pfx + 'Y_W.bias').double()
            Wv, bv = self.p(pfx + 'X_V.weight').double(), self.p(pfx + 'X_V.bias').double()
            H = Wv.shape[0]
            return [(W[:, H:] @ Wv).float().contiguous(), (b + W[:, H:] @ bv).float().contiguous()]
        return [self.derived(('ywfold', pfx, i), lambda i=i: make()[i]) for i in range(2)]

    def a2f(self, pfx, bc, action, rows, rlen, pos_idx, tag, last=True):
        """X2Y_map with X = tokens, Y = rows (frames or segments). -> rows [B,slot,F] act dtype,
        attn_logit [B,slot,Mp], attn [B,slot,Mp] (row = y, col = token)."""
        B, slot, M, F, H = self.B, self.slot, self.ntok, bc['f_dim'], bc['hid_dim']
        Mp = _round_up(M, 4)
        qpos = self.qpos()
        alpha = 1.0 / math.sqrt(H)
        tc = self.x2y_tc(rows, H)
        kt = self.buf('x2y_qt16' if tc else 'x2y_qt', (B, M, H), torch.bfloat16 if tc else torch.float32)   # alpha * Wq^T xk
        cb = self.buf('x2y_c', (B, M, 1))                                   # alpha * xk . bq
        fold = tc and H % 32 == 0
        if fold:     # xk = X_K(action + pos) only feeds linear maps (see f2a)
            Wf, bf_, u, c0 = self._x2y_fold(pfx, 'X_K', 'Y_Q', alpha)
            self.lin(action, Wf, H, kt, pos=qpos, bias=bf_)
            self.lin(action, u, 1, cb, pos=qpos, bias=c0)
        else:
            xk = self.buf('x2y_tokH', (B, M, H))
            self.lin(action, self.p(pfx + 'X_K.weight'), H, xk, pos=qpos, bias=self.p(pfx + 'X_K.bias'))
            self.lin(xk, self.tr(pfx + 'Y_Q.weight'), H, kt, alpha=alpha)
            self.lin(xk, self.p(pfx + 'Y_Q.bias')[None, :], 1, cb, alpha=alpha)
        if (fold and self.frame_pos is None and self.use_fused_x2y and rows.is_contiguous() and ops.a2f_fused_ok(M, H, F, slot)):
            # ONE tcgen05 kernel: logits -> softmax in registers -> value GEMM + Y_W, the rows read once (x2y_fused.cu).  The fp32
            # logits / attention leave the SM only when something reads them: the loss and keep_attn (self.keep), and the
            # attention of the LAST block for the eval fusion.
            Kp = _round_up(M, 64)
            logit = self.buf('a2f_logit_' + tag, (B, slot, Mp)) if self.keep else None
            attn = self.buf('a2f_attn_' + tag, (B, slot, Mp)) if (self.keep or last) else None
            vt = self.zbuf('x2y_vt16', (B, F, Kp), torch.bfloat16)          # pad cols stay 0
            Wav, b2 = self._yw_fold(pfx)
            wa = self.derived(('wav_rep', pfx, B), lambda: Wav.unsqueeze(0).expand(B, -1, -1))
            ops.gemm([S(wa, action)], M, vt, tc=True, tag='x2y_vt')
            out = self.zbuf('a2f_out', (B, slot, F), self.act)
            ops.a2f_fused(rows, kt, cb[:, :, 0], self.wbf(self.p(pfx + 'Y_W.weight')[:, :H]), vt, b2, out, M, logit=logit, attn=attn, len=rlen)
            return out, logit, attn
        logit = self.buf('a2f_logit_' + tag, (B, slot, Mp))
        if tc:      # (the positioned rows of this block were built by f2a a moment ago when fpos is on: same buffer, same values)
            ops.gemm([S(self.with_pos(rows, rlen, pos_idx, 'x2y_rows_pos'), kt)], M, logit, len=rlen, bias=cb[:, :, 0], tc=True, tag='x2y_rows')
        else:
            ops.gemm([S(rows, kt, pos=self.frame_pos, pos_idx=pos_idx)], M, logit, len=rlen, bias=cb[:, :, 0], tag='x2y_rows')
        attn = self.buf('a2f_attn_' + tag, (B, slot, Mp))
        Kp = _round_up(M, 64)
        attn16 = self.buf('a2f_attn16', (B, slot, Kp), torch.bfloat16) if tc else None
        ops.row_softmax(logit, attn, M, len=rlen, out16=attn16)
        W = self.p(pfx + 'Y_W.weight')                                      # [F, 2H] = [Wy | Wa]
        out = self.zbuf('a2f_out', (B, slot, F), self.act)
        if fold:
            # vt[b,f,m] = sum_h Wa[f,h] X_V(action)[b,m,h] = (Wa Wv) action^T + Wa bv; the attention rows sum to one, so the
            # constant Wa bv moves into the output bias.  tf32 tensor-core GEMM: rows = the (replicated) folded weight,
            # per-video "weights" = the tokens themselves.
            vt = self.zbuf('x2y_vt16', (B, F, Kp), torch.bfloat16)          # pad cols stay 0
            Wav, b2 = self._yw_fold(pfx)
            wa = self.derived(('wav_rep', pfx, B), lambda: Wav.unsqueeze(0).expand(B, -1, -1))
            ops.gemm([S(wa, action)], M, vt, tc=True, tag='x2y_vt')
            ops.gemm([S(rows, self.wbf(W[:, :H])), S(attn16, vt)], F, out, len=rlen, bias=b2, tc=True, tag='x2y_rows')
            return out, logit, attn
        xv = self.buf('x2y_xv', (B, M, H))
        self.lin(action, self.p(pfx + 'X_V.weight'), H, xv, bias=self.p(pfx + 'X_V.bias'))
        if tc:
            vt = self.zbuf('x2y_vt16', (B, F, Kp), torch.bfloat16)          # vt[b,f,m] = sum_h Wa[f,h] xv[b,m,h]; pad cols stay 0
            ops.gemm([S(W[None, :, H:], xv)], M, vt)
            ops.gemm([S(rows, self.wbf(W[:, :H])), S(attn16, vt)], F, out, len=rlen, bias=self.p(pfx + 'Y_W.bias'), tc=True, tag='x2y_rows')
        else:
            vt = self.buf('x2y_vt', (B, F, Mp))
            ops.gemm([S(W[None, :, H:], xv)], M, vt)
            ops.gemm([S(rows, W[:, :H]), S(attn, vt, K=M)], F, out, len=rlen, bias=self.p(pfx + 'Y_W.bias'))
        return out, logit, attn

    # ------------------------------------------------------------------ blocks
    def input_block(self, i, bc, x, st):
        pfx = f'block_list.{i}.'
        frame, st['frame_clogit'], st['pred'] = self.frame_branch(pfx + 'frame_branch.', bc, x, True, f'b{i}')
        action = self.action_branch(pfx + 'action_branch.', bc, None, f'b{i}', frame=frame)
        st['action_clogit'] = self.token_splice(action, f'b{i}')
        return frame, action

    def update_block(self, i, bc, frame, action, st):
        pfx, tag = f'block_list.{i}.', f'b{i}'
        tok, st['f2a_attn_logit'], st['f2a_attn'] = self.f2a(pfx + 'f2a_layer.', bc, frame, self.len, None, action, tag, self.keep)
        action = self.action_branch(pfx + 'action_branch.', bc, tok, tag)
        st['action_clogit'] = self.token_splice(action, tag)
        fr, st['a2f_attn_logit'], st['a2f_attn'] = self.a2f(pfx + 'a2f_layer.', bc, action, frame, self.len, None, tag,
                                                            last=(i == len(self.hp['blocks']) - 1))
        frame, st['frame_clogit'], st['pred'] = self.frame_branch(pfx + 'frame_branch.', bc, fr, False, tag)
        return frame, action

    def downsample(self, i, bc, frame, pred, st, gru_layers):
        """temporal_downsample (models/blocks.py:417-437; blocks_SepVerbNoun.py:279-303) with the segmentation kept on device:
        run-length segments of ``pred``, segment mean, (stacked) bi-GRU, ReLU, seg_combine, process_feature.
        Returns (seg rows [B,slot,H], segment count [B], position index or None)."""
        pfx, tag = f'block_list.{i}.', f'b{i}'
        B, slot, H = self.B, self.slot, bc['hid_dim']
        Hh = H // 2
        I32 = torch.int32
        seg_label, seg_start = self.buf('seg_label_' + tag, (B, slot), I32), self.buf('seg_start_' + tag, (B, slot), I32)
        seg_len, seg_center = self.buf('seg_len_' + tag, (B, slot), I32), self.buf('seg_center_' + tag, (B, slot), I32)
        nseg = self.buf('nseg_' + tag, (B,), I32)
        ops.tdu_segment(pred, seg_label, seg_start, seg_len, seg_center, nseg, len=self.len)
        st.update(seg_label=seg_label, seg_lens=seg_len, nseg=nseg, tdu_pred=pred)
        seg0 = self.buf('seg0', (B, slot, H), self.act)
        ws = self.buf('segmean_ws', (ops.segment_mean_ws(B, slot, H),))
        ops.segment_mean(frame, seg0, seg_label, seg_start, seg_len, nseg, ws=ws)
        g = pfx + 'seg_update.'
        cur = seg0
        for l in range(gru_layers):       # layer l > 0 reads the concatenated directions of layer l-1 (nn.GRU stacking)
            gi = self.buf('gru_gi', (B, slot, 6 * Hh))
            self.mm([S(cur, self.cat(f'{g}weight_ih_l{l}', f'{g}weight_ih_l{l}_reverse'))], 6 * Hh, gi, len=nseg,
                    bias=self.cat(f'{g}bias_ih_l{l}', f'{g}bias_ih_l{l}_reverse'), tag='gru_in')
            nxt = self.buf(f'seg1_{l % 2}', (B, slot, H), self.act)
            mma = self.mode == 'bf16' and self.use_tc and Hh == 256
            ops.gru_bidir(gi, self.p(f'{g}weight_hh_l{l}'), self.p(f'{g}bias_hh_l{l}'), self.p(f'{g}weight_hh_l{l}_reverse'),
                          self.p(f'{g}bias_hh_l{l}_reverse'), nxt, nseg, relu=(l == gru_layers - 1), mma=mma,
                          order_ws=self.buf('gru_order', (B,), I32) if (mma and B > 8) else None)
            cur = nxt
        seg2 = self.buf('seg2_' + tag if self.vn is not None else 'seg2', (B, slot, H), self.act)
        self.mm([S(cur, self.p(pfx + 'seg_combine.weight'))], H, seg2, len=nseg, bias=self.p(pfx + 'seg_combine.bias'), tag='seg_combine')
        st['seg_clogit'] = self.buf('seg_clogit_' + tag, (B, slot, self.ncls()))
        self.splice(seg2, st['seg_clogit'], len=nseg)
        return seg2, nseg, (seg_center if self.frame_pos is not None else None)

    def input_block_tdu(self, i, bc, x, st, forced_pred=None):
        """InputBlockTDU.forward (blocks_SepVerbNoun.py:367-398): the tokens attend the SEGMENTS of the frame branch's own
        prediction; the frame feature leaves the block as the frame branch produced it."""
        pfx, tag = f'block_list.{i}.', f'b{i}'
        frame, st['frame_clogit'], st['pred'] = self.frame_branch(pfx + 'frame_branch.', bc, x, True, tag)
        seg, nseg, pidx = self.downsample(i, bc, frame, st['pred'] if forced_pred is None else forced_pred, st, 2)
        action = self.action_branch(pfx + 'action_branch.', bc, None, tag, frame=seg, rlen=nseg, pos_idx=pidx)
        st['action_clogit'] = self.token_splice(action, tag)
        return frame, action

    def update_block_tdu(self, i, bc, frame, action, pred, st):
        """UpdateBlockTDU.forward (models/blocks.py:417-485; blocks_SepVerbNoun.py:445-483)."""
        pfx, tag = f'block_list.{i}.', f'b{i}'
        B, slot, H, F = self.B, self.slot, bc['hid_dim'], bc['f_dim']
        seg2, nseg, pidx = self.downsample(i, bc, frame, pred, st, self.hp['s_layers'])
        seg_label = st['seg_label']
        tok, st['f2a_attn_logit'], st['f2a_attn_seg'] = self.f2a(pfx + 'f2a_layer.', bc, seg2, nseg, pidx, action, tag, self.keep)
        action = self.action_branch(pfx + 'action_branch.', bc, tok, tag)
        st['action_clogit'] = self.token_splice(action, tag)
        seg3, st['a2f_attn_logit'], st['a2f_attn_seg'] = self.a2f(pfx + 'a2f_layer.', bc, action, seg2, nseg, pidx, tag,
                                                                last=(i == len(self.hp['blocks']) - 1))
        W = self.p(pfx + 'sf_merge.0.weight')                               # [F, F+H], input = cat[s2f, frame]
        fr = self.zbuf('sf_out', (B, slot, F), self.act)
        # cat[s2f, frame] W^T = (seg3 W1^T)[seg_label] + frame W2^T: the gather moves to a segment-level product
        s2f = self.buf('sf_pre', (B, slot, F))
        self.mm([S(seg3, W[:, :F])], F, s2f, len=nseg, tag='sf_merge')
        self.mm([S(frame, W[:, F:])], F, fr, len=self.len, bias=self.p(pfx + 'sf_merge.0.bias'), relu=True,
                pre=s2f, pre_idx=seg_label, tag='sf_merge')
        frame, st['frame_clogit'], st['pred'] = self.frame_branch(pfx + 'frame_branch.', bc, fr, False, tag)
        return frame, action

    # ------------------------------------------------------------------ whole forward
    def qpos(self):
        """Query position of the action tokens: the learned action_query, or none for FACT.trans models (blocks.py:79 sets
        it to zero there; the tokens themselves are embed(transcript) + positional encoding)."""
        return None if self.hp['trans'] else self.p('action_query')[:, 0]

    def _feature_dtype(self, seqs):
        """fp32 features (the reference's format, dataset.py:12-21) or, in bf16 mode, features already stored as bf16: half the
        host->device bytes (input staging) and the input projection runs as a bf16 tensor-core GEMM."""
        dts = {s.dtype for s in seqs}
        if dts == {torch.float32}:
            return torch.float32
        if dts == {torch.bfloat16} and self.mode == 'bf16':
            return torch.bfloat16
        raise TypeError(f'features must all be float32 (or all bfloat16 in bf16 compute mode), got {sorted(map(str, dts))}')

    @torch.no_grad()
    def run(self, seqs, forced_preds=None, keep=False, transcript=None):
        """seqs: list of (T_i, in_dim) fp32 tensors, CUDA or (pinned) host.  Synchronous with respect to the stream.
        transcript (FACT.trans models, one video per call): int32 CUDA tensor [N] of the video's action sequence."""
        self._refresh_weights()
        self.ntok, self.action_init, self.transcript = self.hp['ntoken'], None, None
        lengths = [int(s.shape[0]) for s in seqs]
        B, slot, D = len(seqs), _round_up(max(lengths), 128), self.hp['in_dim']
        if self.hp['trans']:
            assert transcript is not None and len(seqs) == 1, 'FACT.trans: one video per call, with its transcript'
            self.ntok = int(transcript.numel())
        self._set_arena((B, slot, self.ntok, self.lane))
        if self.hp['trans']:
            N, A = int(transcript.numel()), self.hp['blocks'][0]['a_dim']
            self.ntok, self.transcript = N, transcript.to(torch.int32).contiguous()
            n_pe = 1000 if N <= 1000 else N + 10        # basic.py:125-127: the table regrows past max_len
            pe = self.derived(('tok_pe', A, n_pe), lambda: _pos_table(A, n_pe, self.dev))
            self.action_init = self.buf('tok_init', (1, N, A))
            if self.vn is None:
                ops.embed_tokens(self.p('action_embed.weight'), self.transcript, pe, self.action_init)
            else:       # verb/noun model (blocks_SepVerbNoun.py:74-83): halves from the verb and the noun of every action
                vids, nids = self.vn_table()
                tr, h = self.transcript.long(), A // 2
                ops.embed_tokens(self.p('verb_embed.weight'), vids[tr].contiguous(), pe, self.action_init[:, :, :h])
                ops.embed_tokens(self.p('noun_embed.weight'), nids[tr].contiguous(), pe[:, h:], self.action_init[:, :, h:])
        x = self.buf('input', (B, slot, D), self._feature_dtype(seqs))
        for b, s in enumerate(seqs):
            x[b, :lengths[b]].copy_(s, non_blocking=True)
        ln = self.buf('len', (B,), torch.int32)
        ln.copy_(torch.tensor(lengths, dtype=torch.int32), non_blocking=True)
        return self.run_packed(x, ln, lengths, forced_preds, keep)

    @torch.no_grad()
    def submit(self, seqs, channel_major=False):
        """Pipelined forward: the host->device copy of this batch runs on a side stream into one of two input
        buffers and overlaps the kernels of the previously submitted batch; the predictions are copied to pinned
        host memory asynchronously.  Returns a handle; ``handle.result()`` blocks until this batch is done.
        At most TWO batches are in flight: submitting a third waits for the oldest one and keeps its predictions (the
        two device / pinned-host slots alternate), so a handle's ``result()`` is valid whenever it is called.
        channel_major: ``seqs[i]`` is the (D, T_i) fp32 array as stored on disk (what the reference transposes on the host,
        utils/dataset.py:12-21); it is copied as it is and transposed on the device, on the copy stream."""
        self._refresh_weights()
        if self.hp['trans']:
            raise NotImplementedError('FACT.trans models run one video per call through forward() (the transcript sets the token count)')
        self.ntok, self.action_init, self.transcript = self.hp['ntoken'], None, None
        lengths = [int(s.shape[1 if channel_major else 0]) for s in seqs]
        B, slot, D = len(seqs), _round_up(max(lengths), 128), self.hp['in_dim']
        k = self._submits % 2
        self._submits += 1
        old = self._slot_pending[k]
        if old is not None:                 # the batch that last used this slot: take its predictions out of the pinned buffer
            old.detach_result()
        self._set_arena((B, slot, self.ntok, self.lane))
        main = torch.cuda.current_stream()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        if channel_major:
            assert all(s.dtype == torch.float32 and s.shape[0] == D and s.is_contiguous() for s in seqs), \
                'channel-major staging takes contiguous (D, T) float32 arrays'
            x = self.buf(f'input_p{k}', (B, slot, D), self.act)        # bf16 rows in bf16 mode: the cast rides on the transpose
            raw = self.buf(f'input_cm_p{k}', (B, D * slot))
        else:
            x = self.buf(f'input_p{k}', (B, slot, D), self._feature_dtype(seqs))
        ln = self.buf(f'len_p{k}', (B,), torch.int32)
        pred = self.buf(f'pred64_p{k}', (B, slot), torch.int64)
        key = (f'pred_host_p{k}', B, slot)
        if key not in self._bufs:
            self._bufs[key] = torch.empty(B, slot, dtype=torch.int64, pin_memory=True)
        host = self._bufs[key]
        free = self._slot_free[k]           # (an event of ANY arena: input buffers of different shapes do not alias, waiting is harmless)
        with torch.cuda.stream(self._copy_stream):
            if free is not None:
                self._copy_stream.wait_event(free)          # the previous user of this input buffer has finished
            ln.copy_(torch.tensor(lengths, dtype=torch.int32).pin_memory(), non_blocking=True)
            if channel_major:
                for b, s in enumerate(seqs):
                    raw[b, :D * lengths[b]].copy_(s.reshape(-1), non_blocking=True)
                ops.transpose_rows(raw, raw.stride(0), x, ln, D)
            else:
                for b, s in enumerate(seqs):
                    x[b, :lengths[b]].copy_(s, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self._copy_stream)
        main.wait_event(copied)
        out = self.run_packed_graphed(x, ln, lengths, pred_out=pred)
        host.copy_(pred, non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        self._slot_free[k] = done
        h = self._slot_pending[k] = _Pending(done, host, lengths, out)
        return h

    def _vn_eval(self, out, pred_out):
        """Verb/noun model: action log-probabilities (combine_verb_noun_to_action, blocks_SepVerbNoun.py:188-226) of the last
        block -- of every block when the attributes are kept -- and Block._eval on them (:307-329)."""
        B, slot, M, ln = self.B, self.slot, self.ntok, self.len
        n1, n2 = self.hp['n_classes']
        vids, nids = self.vn_table()
        A = vids.numel()
        for i, st in enumerate(out['blocks']):
            if not (self.keep or st is out['blocks'][-1]):
                continue
            tag = f'b{i}'
            st['frame_logp'] = self.buf('frame_logp_' + tag, (B, slot, A))
            ops.vn_combine(st['frame_clogit'], n1, n2, vids, nids, st['frame_logp'], len=ln)
            st['action_logp'] = self.buf('action_logp_' + tag, (B, M, A + 1))
            ops.vn_combine(st['action_clogit'], n1 + 1, n2 + 1, vids, nids, st['action_logp'], with_null=True)
            if self.keep:
                st['seg_logp'] = self.buf('seg_logp_' + tag, (B, slot, A))
                ops.vn_combine(st['seg_clogit'], n1, n2, vids, nids, st['seg_logp'], len=st['nseg'])
        last = out['blocks'][-1]
        assert 'a2f_attn_seg' in last, 'the verb/noun model ends in an update block (its eval reads a2f_attn)'
        pred64 = pred_out if pred_out is not None else self.buf('pred64', (B, slot), torch.int64)
        if self.hp['trans']:
            # _eval_w_transcript (blocks_SepVerbNoun.py:331-336): the transcript entry whose token attends the frame most --
            # the transcript fusion kernel with a zero frame weight (softmax keeps the arg max of the attention)
            ntr = self.buf('ntr', (B,), torch.int32)
            ntr.fill_(self.ntok)
            ops.fuse_eval_transcript(last['a2f_attn_seg'], last['frame_logp'], 0.0, self.transcript[None], ntr, pred64, A,
                                     seg_label=last['seg_label'], len=ln)
        else:
            ops.fuse_eval(last['action_logp'], last['a2f_attn_seg'], last['frame_logp'], self.hp['mwt'], pred64, M, A,
                          seg_label=last['seg_label'], len=ln, f_logp=True)
        out['pred'] = pred64
        return out

    @torch.no_grad()
    def run_packed_graphed(self, x, ln, lengths, pred_out=None):
        """``run_packed`` captured once per (input buffer, lengths) into a CUDA graph and replayed afterwards: the forward has
        no host synchronisation and only touches cached buffers, so the ~240 kernel launches of a batch cost one graph
        launch.  Kernel timers (bench --profile) and ``use_graph = False`` fall back to eager launches."""
        if not self.use_graph or ops.TIMER is not None:
            return self.run_packed(x, ln, lengths, pred_out=pred_out)
        self._refresh_weights()
        self._set_arena((x.shape[0], x.shape[1], self.ntok, self.lane))
        # the captured launches depend on B, slot and the buffer pointers only: the lengths reach the kernels through the
        # device tensor ``ln``, so batches of the same shape with other lengths replay the same graph
        key = (x.data_ptr(), tuple(x.shape), x.dtype, ln.data_ptr(), None if pred_out is None else pred_out.data_ptr())
        ent = self._graphs.get(key)
        if ent is None:
            if not self._arena['warmed']:       # eager once per arena: allocates its buffers, fills the weight caches
                self.run_packed(x, ln, lengths, pred_out=pred_out)
                self._arena['warmed'] = True
            elif self._len_sig != tuple(lengths):
                for t in self._zbufs.values():
                    t.zero_()
                self._len_sig = tuple(lengths)
            n0 = ops.COUNTERS['launches']
            nb = len(self._bufs)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.run_packed(x, ln, lengths, pred_out=pred_out)
            assert len(self._bufs) == nb, 'a buffer was allocated during graph capture (it would live in the graph pool)'
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            ent = self._graphs[key] = (g, out, ops.COUNTERS['launches'] - n0)
            self.last_launches = ent[2]
            g.replay()          # the capture itself did not execute
            return out
        if self._len_sig != tuple(lengths):      # other lengths than the last run of this arena: restore the zero tails the taps rely on
            for t in self._zbufs.values():
                t.zero_()
            self._len_sig = tuple(lengths)
        ent[0].replay()
        ops.COUNTERS['launches'] += ent[2]
        self.last_launches = ent[2]
        ent[1]['lengths'] = lengths
        return ent[1]

    @torch.no_grad()
    def run_packed(self, x, ln, lengths, forced_preds=None, keep=False, pred_out=None):
        self._refresh_weights()
        hp = self.hp
        self._set_arena((x.shape[0], x.shape[1], self.ntok, self.lane))
        self.B, self.slot, self.len, self.keep = x.shape[0], x.shape[1], ln, keep
        if self._len_sig != tuple(lengths):
            for t in self._zbufs.values():
                t.zero_()
            self._len_sig = tuple(lengths)
        B, slot, M, C, H = self.B, self.slot, self.ntok, hp['n_classes'], hp['blocks'][0]['hid_dim']
        self.frame_pos = None
        if hp['fpos']:
            self.frame_pos = self.derived(('pe', slot, H), lambda: _pos_table(H, slot, self.dev))
        frame, action, stash, u = x, None, [], 0
        pred = None
        for i, bc in enumerate(hp['blocks']):
            st = {}
            if bc['type'] == 'i':
                frame, action = self.input_block(i, bc, frame, st)
            elif bc['type'] == 'I':
                fp = None
                if forced_preds is not None:
                    fp = self.buf(f'forced_pred_{u}', (B, slot), torch.int32)
                    for b in range(B):
                        fp[b, :lengths[b]].copy_(forced_preds[u][b].to(torch.int32), non_blocking=True)
                frame, action = self.input_block_tdu(i, bc, frame, st, fp)
                u += 1
            elif bc['type'] == 'u':
                frame, action = self.update_block(i, bc, frame, action, st)
            else:
                if forced_preds is not None:
                    fp = self.buf(f'forced_pred_{u}', (B, slot), torch.int32)
                    for b in range(B):
                        fp[b, :lengths[b]].copy_(forced_preds[u][b].to(torch.int32), non_blocking=True)
                    pred = fp
                frame, action = self.update_block_tdu(i, bc, frame, action, pred, st)
                u += 1
            pred = st['pred']
            st['frame_feature'], st['action_feature'] = frame, action
            stash.append(st)
        last = stash[-1]
        out = dict(blocks=stash, lengths=lengths)
        if self.vn is not None:
            return self._vn_eval(out, pred_out)
        if self.clip and 'text_embeddings' in self._p and self._p['text_embeddings'] is not None:
            P = self.p('frame_projection.projection.0.weight').shape[0]
            h1 = self.buf('clip_h1', (B, slot, P), self.act)
            w0 = self.derived(('clip_w0pad',), lambda: torch.nn.functional.pad(self.p('frame_projection.projection.0.weight'), (0, C)))
            self.mm([S(frame, w0)], P, h1, len=ln, tag='clip',
                    bias=self.p('frame_projection.projection.0.bias'))
            ops.layernorm(h1, self.p('frame_projection.projection.1.weight'), self.p('frame_projection.projection.1.bias'),
                          h1, relu=True, len=ln)
            # bf16 mode: the embedding stays bf16 end to end (CTA-pair GEMM -> in-place L2 normalise -> bf16 text GEMM): half the
            # bytes of the fp32 round trip the three kernels used to make (1 GB per 64 x 4096 frames for the normalisation alone)
            emb = self.buf('clip_emb', (B, slot, 512), self.act)
            self.mm([S(h1, self.p('frame_projection.projection.4.weight'))], 512, emb, len=ln, tag='clip',
                    bias=self.p('frame_projection.projection.4.bias'))
            ops.l2norm(emb, emb, len=ln)
            flogit = self.buf('clip_logit', (B, slot, C))
            self.mm([S(emb, self.p('text_embeddings'))], C, flogit, tf32=True, len=ln, alpha=1.0 / hp['temp'], tag='clip')
            out['projected_frame_embeddings'], out['clip_logit'] = emb, flogit
        else:
            flogit = last['frame_clogit']
        pred64 = pred_out if pred_out is not None else self.buf('pred64', (B, slot), torch.int64)
        if self.hp['trans'] and 'clip_logit' not in out:
            # Block._eval_w_transcript (blocks.py:263-275); FACT_CLIP with text embeddings keeps the token / CLIP fusion
            ntr = self.buf('ntr', (B,), torch.int32)
            ntr.fill_(self.ntok)
            seg = last.get('a2f_attn_seg')
            ops.fuse_eval_transcript(seg if seg is not None else last['a2f_attn'], last['frame_clogit'], hp['mwt'],
                                     self.transcript[None], ntr, pred64, C,
                                     seg_label=last['seg_label'] if seg is not None else None, len=ln)
        elif 'a2f_attn' in last:
            ops.fuse_eval(last['action_clogit'], last['a2f_attn'], flogit, hp['mwt'], pred64, M, C, len=ln)
        elif 'a2f_attn_seg' in last:
            ops.fuse_eval(last['action_clogit'], last['a2f_attn_seg'], flogit, hp['mwt'], pred64, M, C,
                          seg_label=last['seg_label'], len=ln)
        else:
            ops.fuse_eval(None, None, flogit, hp['mwt'], pred64, 0, C, len=ln)
        out['pred'] = pred64
        return out


class _Pending:
    """Handle of a submitted batch (FactEngine.submit)."""

    def __init__(self, done, host_pred, lengths, out):
        self.done, self.host_pred, self.lengths, self.out = done, host_pred, lengths, out

        self._res = None

    def detach_result(self):
        """Copy the predictions out of the shared pinned buffer (called before the buffer's slot is reused)."""
        if self._res is None:
            self.done.synchronize()
            p = self.host_pred.numpy()
            self._res = [{'pred': p[b, :T].copy()} for b, T in enumerate(self.lengths)]
            self.host_pred = None
        return self._res

    def result(self):
        return self.detach_result()


def _pos_table(d_model, length, device):
    """Sinusoid table of PositionalEncoding (models/basic.py:90-102), built on the host once per length."""
    pe = torch.zeros(length, d_model)
    position = torch.arange(0, length, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(device)
