"""Training step of FACT / FACT_CLIP on the factk kernels: train-mode forward, loss, hand-written backward.

The reference trains with ``loss, saves = net(seqs, labels, compute_loss=True); loss.backward()`` (scripts/train.py:262-264),
i.e. torch autograd over eager cuDNN / cuBLAS ops, one video at a time.  Here the whole batch runs as ONE train-mode forward
over the same kernels the inference engine uses (GEMMs, segmentation, GRU, softmax-splice ...), every intermediate that a
gradient needs is kept in HBM, and the backward pass is a tape of closures that launch the kernels of csrc/train*.cu:

  * data gradient of a Conv1d tap / Linear  = the forward GEMM with the transposed weight and the negated tap offset
  * weight gradient                         = ``factk_wgrad`` (rows^T x rows, split over the frames, fixed-order reduction)
  * attention (token self-attention, tokens-attend-frames, X2Y f2a / a2f) is spelled out as logit GEMM -> softmax ->
    apply, so its backward is the same three primitives again (no attention-specific gradient kernel)
  * bi-GRU: back-propagation through time on a 4-CTA cluster per (video, direction) (``factk_gru_bwd``)
  * LayerNorm, softmax-splice (process_feature), L2-normalise, segment mean / gather: one row kernel each
  * the loss gradients w.r.t. every logit tensor: csrc/train_loss.cu

Derived weights (Conv1d taps re-laid as [tap][out][in], the folded MSTCN++ layer, concatenated GRU input weights, the padded
CLIP projection) are differentiable torch functions of the parameters: the kernels produce the gradient of the DERIVED
tensor and one ``torch.autograd.grad`` over those parameter-sized functions carries it to the parameters.  Nothing
activation-sized ever goes through torch autograd.

Train-mode augmentations of the reference forward (blocks.py:614-622): ``Dropout2d`` channel masking of the input features
(FACT.cmr), ``time_mask`` (TM.*), and the dropouts inside the network (cfg dropout, CLIP.projection_dropout).  Masks come from a
counter-based hash of (seed, site, element) so the backward pass regenerates them; ``engine.seed`` changes every step.
"""
import math
import os
import random

import torch

from . import ops
from .engine import FactEngine, _pos_table, _round_up
from .ops import S


class Var:
    """One activation of the training forward: value, lazily allocated gradient, valid-row vector."""
    __slots__ = ('v', 'g', 'len', 'needs_grad')

    def __init__(self, v, ln=None, g=None, needs_grad=True):
        self.v, self.g, self.len, self.needs_grad = v, g, ln, needs_grad


class Wt:
    """A weight as the kernels see it: value ``w`` and gradient accumulator ``g`` (same shape; views slice both)."""
    __slots__ = ('w', 'g', '_t', '_par')

    def __init__(self, w, g):
        self.w, self.g, self._t, self._par = w, g, None, None

    def __getitem__(self, idx):
        c = Wt(self.w[idx], None if self.g is None else self.g[idx])
        if isinstance(idx, int) and self.w.dim() > 2:
            c._par = (self, idx)          # a tap of a stacked weight: its transpose is a slice of the stack's (one copy kernel)
        return c

    def T(self):
        """Transposed (last two dimensions) dense copy, made once per step."""
        if self._t is None:
            self._t = self._par[0].T()[self._par[1]] if self._par is not None else dense(self.w.transpose(-1, -2))
        return self._t


def dense(t):
    """Copy with standard row-major strides (``contiguous()`` keeps odd strides on size-1 dimensions; the kernels take the
    row stride as the leading dimension)."""
    out = torch.empty(t.shape, dtype=t.dtype, device=t.device)
    out.copy_(t)
    return out


def src(x, W, off=0, pos=None, pos_idx=None, K=None):
    return dict(x=x, W=W, off=off, pos=pos, pos_idx=pos_idx, K=K)


class TrainEngine(FactEngine):
    def __init__(self, module, hp, clip, mode='fp32'):
        super().__init__(module, hp, clip, mode)
        self.seed, self.step_no = 0x5EED, 0
        self.forced_masks = None          # tests: {'cmr': keep mask [B, D]} injected instead of the hashed one
        self.use_graphs = False           # capture the step into CUDA graphs per batch shape (net.train_graphs = True)
        self._steps, self._cur, self._cap_stream, self._const, self._seed_dev = {}, None, None, {}, None
        # parameter gradients (weight / bias reductions) are off the backward pass's critical path: they run on a second stream and
        # fill the SMs that the latency-bound stretches (GRU chain, token-side kernels) leave idle; joined at every section's end
        self.side_wgrad = os.environ.get('FACTK_SIDE_WGRAD', '1') != '0'
        self.tc_cross_attn = os.environ.get('FACTK_TC_CROSS_ATTN', '1') != '0'      # SCA cross attention as block-diagonal tcgen05 GEMMs
        self._side, self._side_keep, self._side_busy = None, [], False
        self.tape, self._posrows = [], {}
        if hp['trans'] or self.vn is not None:
            raise NotImplementedError('training step: FACT.trans and the Epic verb/noun model are not built (query-token FACT / '
                                      'FACT_CLIP only)')

    # ------------------------------------------------------------------ bookkeeping
    @property
    def seed_dev(self):
        """Device scalar added to every dropout seed (the step number): static address, new value every step."""
        if self._seed_dev is None or self._seed_dev.device != self.dev:
            self._seed_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        return self._seed_dev

    def const_table(self, key, fn):
        """Constant of the model shape (positional table): survives the per-step weight-cache reset."""
        t = self._const.get((key, self.dev))
        if t is None:
            t = self._const[(key, self.dev)] = fn().contiguous()
        return t

    def begin(self):
        self.tape, self._site, self._posrows = [], 0, {}
        self._params = dict(self.m.named_parameters())
        self._derived = []                # (tensor with autograd graph, gradient accumulator) of the current section
        self._wt, self._leaf = {}, {}
        self._wcache = {}                 # bf16 / transposed copies are keyed by address: never let them outlive a step
        self._layout()
        for f in self._flat:
            f.zero_()

    def _layout(self):
        """Gradient buckets: ONE flat fp32 buffer per section of the model -- block i (+ the action queries in block 0), and
        the CLIP projection head -- with every parameter's gradient a view into it.  A section's buffer is final as soon as
        the backward pass has left the section, so a data-parallel all-reduce of it (in place, no packing copy) overlaps the
        backward pass of the earlier sections."""
        sig = tuple((n, tuple(p.shape), p.device) for n, p in self._params.items())
        if getattr(self, '_layout_sig', None) == sig:
            # a parameter still holding last step's gradient view (no zero_grad(set_to_none=True) in between: gradient
            # accumulation): leave those buffers to the parameters and start fresh ones
            if not any(p.grad is not None and p.grad.data_ptr() == self._pg[n].data_ptr() for n, p in self._params.items()):
                return
        nb = len(self.hp['blocks'])
        def section(n):
            if n.startswith('block_list.'):
                return int(n.split('.')[1])
            return nb if n.startswith('frame_projection.') else 0
        names = [[] for _ in range(nb + 1)]
        for n in self._params:
            names[section(n)].append(n)
        self._flat, self._pg, self._sect_names = [], {}, names
        for k in range(nb + 1):
            tot = sum(_round_up(self._params[n].numel(), 4) for n in names[k])
            flat = torch.zeros(max(tot, 1), dtype=torch.float32, device=self.dev)
            off = 0
            for n in names[k]:
                p = self._params[n]
                self._pg[n] = flat[off:off + p.numel()].view(p.shape)
                off += _round_up(p.numel(), 4)
            self._flat.append(flat)
        self._layout_sig = sig

    def mark_section(self, k):
        """Tape marker at the START of section k of the forward pass: when the (reversed) tape reaches it, every gradient
        contribution to the section's parameters has been launched."""
        self._derived = []
        self.tape.append(('section', k, self._derived))

    def _section_done(self, k):
        hook = getattr(self.m, 'grad_ready_hook', None)
        if hook is not None:
            hook(k, self._flat[k], self._sect_names[k])

    def _finish_derived(self, derived):
        """Carry the gradients of a section's derived weights (parameter-sized differentiable functions) to the parameters."""
        outs = [t for t, _ in derived if t.requires_grad]
        if not outs:
            return
        gouts = [g for t, g in derived if t.requires_grad]
        names = [n for n, t in self._leaf.items() if t.requires_grad]
        gs = torch.autograd.grad(outs, [self._leaf[n] for n in names], gouts, allow_unused=True)
        for n, g in zip(names, gs):
            if g is not None:
                self._pg[n] += g

    def new(self, shape, dtype=torch.float32, zero=False):
        return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.dev)

    def G(self, var):
        if var.g is None:
            var.g = torch.zeros_like(var.v)
        return var.g

    def cols(self, var, a, b):
        """Column slice of a Var sharing value and gradient storage."""
        self.G(var)
        return Var(var.v[..., a:b], var.len, var.g[..., a:b])

    def W(self, name):
        """Parameter as a weight handle (gradient accumulates in a buffer of the parameter's shape)."""
        h = self._wt.get(name)
        if h is None:
            h = self._wt[name] = Wt(self._params[name].detach(), self._pg[name])
        return h

    def const(self, t):
        return Wt(t, None)

    def D(self, key, fn):
        """Derived weight: ``fn()`` is a differentiable torch function of the parameters (parameter-sized)."""
        h = self._wt.get(key)
        if h is None:
            with torch.enable_grad():
                t = fn()
            g = torch.zeros(t.shape, dtype=torch.float32, device=t.device)
            self._derived.append((t, g))
            h = self._wt[key] = Wt(dense(t.detach()), g)
        return h

    def Dn(self, key, fn):
        """Several derived weights from one differentiable function (``fn()`` returns a tuple of tensors)."""
        h = self._wt.get(key)
        if h is None:
            with torch.enable_grad():
                ts = fn()
            hs = []
            for t in ts:
                g = torch.zeros(t.shape, dtype=torch.float32, device=t.device)
                self._derived.append((t, g))
                hs.append(Wt(dense(t.detach()), g))
            h = self._wt[key] = tuple(hs)
        return h

    def P(self, name):
        """Parameter as the input of a derived-weight function: a fresh leaf sharing the parameter's storage.  Fresh per step on
        purpose -- the parameter's own AccumulateGrad node outlives a step while anything (a kept ``loss`` tensor) references
        last step's graph, and it remembers the stream it was created on; reaching it from inside a stream capture makes
        autograd synchronise the capturing stream with that (default) stream, which invalidates the capture."""
        t = self._leaf.get(name)
        if t is None:
            t = self._leaf[name] = self._params[name].detach().requires_grad_(self._params[name].requires_grad)
        return t

    def taps_w(self, name, groups=1):
        """Conv1d weight (Cout, Cin/groups, k) as [k][Cout][Cin] (dense block-diagonal for grouped convolutions)."""
        def make():
            w = self.P(name)
            if groups > 1:
                co, ci = w.shape[0] // groups, w.shape[1]
                full = w.new_zeros(w.shape[0], ci * groups, w.shape[2])
                for j in range(groups):
                    full[j * co:(j + 1) * co, j * ci:(j + 1) * ci] = w[j * co:(j + 1) * co]
                w = full
            return w.permute(2, 0, 1).contiguous()
        return self.D(('taps', name, groups), make)

    def next_site(self):
        self._site += 1
        return self._site

    # ------------------------------------------------------------------ primitive ops (forward + tape entry)
    def linear(self, srcs, N, bias=None, relu=False, res=None, pre=None, pre_seg=None, alpha=1.0, ln=None, dtype=None, B=None,
               rows=None, tag=None):
        """y = act(alpha * sum_s (x_s[t + off_s] + pos_s) W_s^T + bias + pre[idx]) + res.  W: Wt (shared) or Var ([B,N,K] per video);
        bias: Wt [N] or Var [B,N]; pre: Var of rows gathered by pre_seg['label'] (a segmentation dict) or taken row for row."""
        assert not (relu and res is not None), 'ReLU and residual in one training GEMM: the mask would need the pre-residual value'
        x0 = srcs[0]['x']
        B = x0.v.shape[0] if B is None else B
        rows = x0.v.shape[1] if rows is None else rows
        dtype = dtype or (self.act if x0.v.dtype == torch.bfloat16 else torch.float32)
        y = Var(self.new((B, rows, N), dtype, zero=(self.mode == 'bf16' and ln is not None)), ln)      # zero tails feed TMA taps
        wv = lambda W: W.v if isinstance(W, Var) else W.w
        ss = [S(s['x'].v, wv(s['W']), K=s['K'], off=s['off'], pos=s['pos'], pos_idx=s['pos_idx']) for s in srcs]
        bv = None if bias is None else (bias.v if isinstance(bias, Var) else bias.w)
        self.mm(ss, N, y.v, len=ln, bias=bv, relu=relu, res=None if res is None else res.v, alpha=alpha,
                pre=None if pre is None else pre.v, pre_idx=None if pre_seg is None else pre_seg['label'], tag=tag)

        def bwd():
            if y.g is None:
                return
            dz = y.g
            if res is not None and res.needs_grad:
                ops.ew(ops.EW_AXPY, dz, self.G(res), N, len=ln)
            if relu:
                ops.ew(ops.EW_RELU_BWD, dz, dz, N, r=y.v, len=ln)
            if bias is not None:
                if isinstance(bias, Var):
                    ops.colsum(dz, N, self.G(bias), len=ln, per_video=True)
                elif bias.g is not None:
                    self.pgrad(lambda: ops.colsum(dz, N, bias.g, len=ln), dz)
            if pre is not None:
                if pre_seg is None:
                    ops.ew(ops.EW_AXPY, dz, self.G(pre), N, len=ln)
                else:
                    ops.segment_reduce(dz, self.G(pre), pre_seg['start'], pre_seg['len'], pre_seg['nseg'], N, mean=False, accumulate=True)
            done = set()
            for i, s in enumerate(srcs):
                x, Wh = s['x'], s['W']
                K = s['K'] if s['K'] is not None else wv(Wh).shape[-1]
                # weight gradient
                if isinstance(Wh, Var):
                    self.wgrad(dz, x.v, N, K, self.G(Wh), off=s['off'], len=ln, alpha=alpha, pos=s['pos'], pos_idx=s['pos_idx'],
                               per_video=True)
                elif Wh.g is not None:
                    self.pgrad(lambda x=x, Wh=Wh, K=K, s=s: self.wgrad(dz, x.v, N, K, Wh.g, off=s['off'], len=ln, alpha=alpha, pos=s['pos'],
                                                                      pos_idx=s['pos_idx']), dz, x.v)
                # data gradient: every tap of the same x in one multi-source GEMM accumulating into x.g
                if not x.needs_grad or i in done:
                    continue
                group = [j for j, s2 in enumerate(srcs) if s2['x'] is x]
                done.update(group)
                gs = []
                for j in group:
                    Wj = srcs[j]['W']
                    Kj = srcs[j]['K'] if srcs[j]['K'] is not None else wv(Wj).shape[-1]
                    assert Kj == K
                    if isinstance(Wj, Var):
                        wt = self.new((Wj.v.shape[0], Kj, N))
                        ops.transpose(Wj.v[:, :, :Kj], wt)
                    else:
                        wt = Wj.T() if Kj == Wj.w.shape[-1] else dense(Wj.w[:, :Kj].t())
                    gs.append(S(dz, wt, off=-srcs[j]['off']))
                xg = self.G(x)
                self.mm(gs, K, xg[..., :K] if xg.shape[-1] != K else xg, len=ln, alpha=alpha, res=xg[..., :K] if xg.shape[-1] != K else xg,
                        tag='dgrad')
        self.tape.append(bwd)
        return y

    def mm(self, srcs, N, out, **kw):
        """GEMM dispatch of the training step.  fp32 mode: the fp32 CUDA-core kernel (1e-3 gradient parity).  bf16 mode: the
        inference engine's dispatch -- tcgen05 CTA-pair / multi-source kernels for bf16 operands, tf32 tcgen05 for fp32
        operands, CUDA cores for what they reject (positional terms on a source, odd widths)."""
        kw = {k: v for k, v in kw.items() if v is not None}
        if self.mode == 'bf16':
            return FactEngine.mm(self, srcs, N, out, tf32=True, **kw)
        return ops.gemm(srcs, N, out, **kw)

    def wgrad(self, dz, a, N, K, dw, **kw):
        return ops.wgrad(dz, a, N, K, dw, tc=(self.mode == 'bf16'), **kw)

    def pgrad(self, fn, *keep):
        """Run ``fn`` (launches that only produce PARAMETER gradients) on the side stream, after everything launched so far.
        ``keep``: the tensors it reads -- held until the join so the allocator cannot hand their memory to the main stream."""
        if not self.side_wgrad:
            return fn()
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            fn()
        self._side_keep.extend(keep)
        self._side_busy = True

    def _side_join(self):
        if self._side_busy:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_busy = False
        self._side_keep.clear()

    def add(self, a, b, N=None):
        N = a.v.shape[-1] if N is None else N
        y = Var(torch.empty_like(a.v), a.len)
        ops.ew(ops.EW_ADD, a.v, y.v, N, r=b.v, len=a.len)

        def bwd():
            if y.g is None:
                return
            for t in (a, b):
                if t.needs_grad:
                    ops.ew(ops.EW_AXPY, y.g, self.G(t), N, len=a.len)
        self.tape.append(bwd)
        return y

    def relu(self, x):
        N = x.v.shape[-1]
        y = Var(torch.empty_like(x.v), x.len)
        ops.ew(ops.EW_RELU, x.v, y.v, N, len=x.len)

        def bwd():
            if y.g is None:
                return
            t = torch.empty_like(y.g)
            ops.ew(ops.EW_RELU_BWD, y.g, t, N, r=y.v, len=x.len)
            ops.ew(ops.EW_AXPY, t, self.G(x), N, len=x.len)
        self.tape.append(bwd)
        return y

    def dropout(self, x, p, channel=False, dtype=None):
        """nn.Dropout (or nn.Dropout2d over whole channels) in training mode; identity when p == 0 (a cast when ``dtype``
        differs from the input's)."""
        dtype = dtype or x.v.dtype
        if p <= 0.0:
            if dtype == x.v.dtype:
                return x
            assert not x.needs_grad
            y = Var(torch.zeros(x.v.shape, dtype=dtype, device=self.dev), x.len, needs_grad=False)
            ops.ew(ops.EW_COPY, x.v, y.v, x.v.shape[-1], len=x.len)
            return y
        N, site = x.v.shape[-1], self.next_site()
        op = ops.EW_DROPOUT_CH if channel else ops.EW_DROPOUT
        y = Var(torch.zeros(x.v.shape, dtype=dtype, device=self.dev), x.len, needs_grad=x.needs_grad)
        forced = None if self.forced_masks is None else self.forced_masks.get(site)
        if forced is not None:          # tests inject the keep mask (already scaled by 1/(1-p)) instead of the hash
            ops.ew(ops.EW_MUL, x.v, y.v, N, r=forced, len=x.len)
        else:
            ops.ew(op, x.v, y.v, N, len=x.len, p=p, seed=self.seed, site=site, seed_ptr=self.seed_dev)

        def bwd():
            if y.g is None or not x.needs_grad:
                return
            first = x.g is None                      # first writer of x.g: the masked gradient goes straight into it
            t = self.G(x) if first else torch.empty_like(y.g)
            if forced is not None:
                ops.ew(ops.EW_MUL, y.g, t, N, r=forced, len=x.len)
            else:
                ops.ew(op, y.g, t, N, len=x.len, p=p, seed=self.seed, site=site, seed_ptr=self.seed_dev)
            if not first:
                ops.ew(ops.EW_AXPY, t, self.G(x), N, len=x.len)
        self.tape.append(bwd)
        return y

    def addpos(self, x, qpos):
        """add_positional_encoding (basic.py:313-320) with a learned table: y = x, y[..., :d] += qpos (d = table width)."""
        if qpos is None:
            return x
        d, H = qpos.w.shape[-1], x.v.shape[-1]
        y = Var(x.v.clone(), x.len)
        ops.ew(ops.EW_AXPY, qpos.w, y.v, d, bcast=True)

        def bwd():
            if y.g is None:
                return
            ops.ew(ops.EW_AXPY, y.g, self.G(x), H)
            if qpos.g is not None:
                B, M = y.g.shape[0], y.g.shape[1]
                if y.g.is_contiguous() and d == H:
                    ops.colsum(y.g.view(1, B, M * H), M * H, qpos.g.view(-1))
                else:
                    tmp = self.new((1, M, H))
                    ops.colsum(y.g.reshape(1, B, M * H), M * H, tmp.view(-1), accumulate=False)
                    ops.ew(ops.EW_AXPY, tmp, qpos.g.view(1, M, d), d)
        self.tape.append(bwd)
        return y

    def with_pos_t(self, rows, rlen, pos_idx=None):
        """rows + frame_pos[t or centre] (add_positional_encoding, basic.py:313-320) materialised once per rows tensor and step
        (FactEngine.with_pos): every GEMM that reads positioned rows -- the SCA key projections, the f2a / a2f logits -- then takes
        a plain source and stays on the tensor-core kernels, forward and backward."""
        if self.frame_pos is None:
            return rows
        key = (id(rows), None if pos_idx is None else pos_idx.data_ptr())
        hit = self._posrows.get(key)
        if hit is not None and hit[0] is rows:
            return hit[1]
        N = rows.v.shape[-1]
        y = Var(torch.zeros_like(rows.v), rlen, needs_grad=rows.needs_grad)
        ops.ew(ops.EW_ADDTAB, rows.v, y.v, N, r=self.frame_pos, len=rlen, ridx=pos_idx)

        def bwd():
            if y.g is not None and rows.needs_grad:
                ops.ew(ops.EW_AXPY, y.g, self.G(rows), N, len=rlen)
        self.tape.append(bwd)
        self._posrows[key] = (rows, y)
        return y

    def layernorm(self, x, wname, bname, res=None, relu=False, ln=None):
        w, b = self.W(wname), self.W(bname)
        E = x.v.shape[-1]
        y = Var(torch.empty_like(x.v), ln)
        ops.layernorm(x.v, w.w, b.w, y.v, res=None if res is None else res.v, relu=relu, len=ln)

        def bwd():
            if y.g is None:
                return
            dv = self.new(x.v.shape, y.g.dtype, zero=True)
            ops.layernorm_bwd(x.v, w.w, b.w, y.g, dv, w.g, b.g, res=None if res is None else res.v, relu=relu, len=ln)
            for t in (x, res):
                if t is not None and t.needs_grad:
                    ops.ew(ops.EW_AXPY, dv, self.G(t), E, len=ln)
        self.tape.append(bwd)
        return y

    def splice(self, x, n, ln=None, want_pred=False):
        """Block.process_feature (blocks.py:195-202) in place on x's storage -> (spliced rows Var, raw logits Var, argmax)."""
        B, rows, H = x.v.shape
        clogit = Var(self.new((B, rows, n)), ln)
        pred = self.new((B, rows), torch.int32) if want_pred else None
        ops.softmax_splice(x.v, n, clogit.v, pred, len=ln)
        y = Var(x.v, ln)

        def bwd():
            if y.g is None and clogit.g is None:
                return
            x.g = self.new(x.v.shape, x.v.dtype, zero=True)
            ops.splice_bwd(y.v, y.g, clogit.g, x.g, H, n, len=ln)
        self.tape.append(bwd)
        return y, clogit, pred

    def l2norm(self, x, ln=None):
        y = Var(torch.empty_like(x.v), ln)
        ops.l2norm(x.v, y.v, len=ln)

        def bwd():
            if y.g is None:
                return
            x.g = self.new(x.v.shape, x.v.dtype, zero=True)
            ops.l2norm_bwd(x.v, y.g, x.g, len=ln)
        self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------ attention
    def mha_self(self, q, k, v, nhead, p_drop=0.0):
        """Token self-attention core of nn.MultiheadAttention (basic.py:437,500): logits -> row softmax -> (attention
        dropout) -> apply, each product one head-batched launch (ops.heads_mm); q, k, v: Vars [B, M, A]."""
        B, M, A = q.v.shape
        dh, Mp = A // nhead, _round_up(M, 4)
        alpha = 1.0 / math.sqrt(dh)
        mm = lambda a, b, c, m, n, kd, hs, **kw: ops.heads_mm(a, b, c, m, n, kd, nhead, *hs, **kw)
        L_ = self.new((B, M, nhead * Mp), zero=True)
        mm(q.v, k.v, L_, M, M, dh, (dh, dh, Mp), alpha=alpha)
        Pv = Var(self.new((B, M, nhead * Mp), zero=True))
        Pr = Pv.v.view(B, M * nhead, Mp)
        ops.row_softmax(L_.view(B, M * nhead, Mp), Pr, M)
        del L_

        def bwd_softmax():                       # P -> logits -> q, k
            if Pv.g is None:
                return
            dL = self.new((B, M, nhead * Mp), zero=True)
            ops.row_softmax_bwd(Pr, Pv.g.view(B, M * nhead, Mp), dL.view(B, M * nhead, Mp), M)
            mm(dL, k.v, self.G(q), M, dh, M, (Mp, dh, dh), b_kmajor=True, alpha=alpha, accumulate=True)
            mm(dL, q.v, self.G(k), M, dh, M, (Mp, dh, dh), a_kmajor=True, b_kmajor=True, alpha=alpha, accumulate=True)
        self.tape.append(bwd_softmax)
        Pd = self.dropout(Pv, p_drop)            # nn.MultiheadAttention drops attention weights (its own tape entry)
        o = Var(self.new((B, M, A)))
        mm(Pd.v, v.v, o.v, M, dh, M, (Mp, dh, dh), b_kmajor=True)

        def bwd_apply():                         # o -> P, v
            if o.g is None:
                return
            mm(o.g, v.v, self.G(Pd), M, M, dh, (dh, dh, Mp), accumulate=True)
            mm(Pd.v, o.g, self.G(v), M, dh, M, (Mp, dh, dh), a_kmajor=True, b_kmajor=True, accumulate=True)
        self.tape.append(bwd_apply)
        return o

    def cross_attn(self, q, kk, vv, nhead, rlen, p_drop=0.0):
        """Tokens attend rows (SCALayer cross attention core, basic.py:507-514): softmax over the valid rows per (head, token).
        q: Var [B, M, A] (projected queries); kk, vv: Vars [B, slot, A] (projected keys / values).  The probabilities live
        frame-major, [B, slot, nhead * Mp]; the six products are head-batched launches (ops.heads_mm), the reductions over the
        frames (apply, dq) split the frames across CTAs."""
        B, M, A = q.v.shape
        slot = kk.v.shape[1]
        dh, Mp = A // nhead, _round_up(M, 4)
        alpha = 1.0 / math.sqrt(dh)
        if (self.mode == 'bf16' and self.use_tc and self.tc_cross_attn and kk.v.dtype == torch.bfloat16 and vv.v.dtype == torch.bfloat16
                and A % 64 == 0 and slot % 64 == 0):
            return self._cross_attn_tc(q, kk, vv, nhead, rlen, p_drop)
        mm = lambda a, b, c, m, n, kd, hs, **kw: ops.heads_mm(a, b, c, m, n, kd, nhead, *hs, len=rlen, **kw)
        # [frames x heads x tokens] fp32 tensors (157 MB at T = 16384, M = 300): not zero-filled -- every reader is bounded by the
        # valid rows and the M columns per head, which every writer covers
        L_ = self.new((B, slot, nhead * Mp))
        mm(kk.v, q.v, L_, slot, M, dh, (dh, dh, Mp), len_mode=1, alpha=alpha)
        Pv = Var(self.new((B, slot, nhead * Mp)), rlen)
        ops.col_softmax(L_, Pv.v, nhead * Mp, len=rlen)
        del L_

        def bwd_softmax():
            if Pv.g is None:
                return
            dL = Pv.g                                   # in place: element-wise in P and dP once the column sums exist
            ops.col_softmax_bwd(Pv.v, Pv.g, dL, nhead * Mp, len=rlen)
            mm(dL, q.v, self.G(kk), slot, dh, M, (Mp, dh, dh), b_kmajor=True, len_mode=1, alpha=alpha, accumulate=True)
            mm(dL, kk.v, self.G(q), M, dh, slot, (Mp, dh, dh), a_kmajor=True, b_kmajor=True, len_mode=2, alpha=alpha, accumulate=True)
        self.tape.append(bwd_softmax)
        Pd = self.dropout(Pv, p_drop)
        o = Var(self.new((B, M, A)))
        mm(Pd.v, vv.v, o.v, M, dh, slot, (Mp, dh, dh), a_kmajor=True, b_kmajor=True, len_mode=2)

        def bwd_apply():
            if o.g is None:
                return
            fresh = Pd.g is None                        # first (only) writer of dP: no zero fill, no read-modify-write
            if fresh:
                Pd.g = self.new(Pd.v.shape)
            mm(vv.v, o.g, Pd.g, slot, M, dh, (dh, dh, Mp), len_mode=1, accumulate=not fresh)
            mm(Pd.v, o.g, self.G(vv), slot, dh, M, (Mp, dh, dh), b_kmajor=True, len_mode=1, accumulate=True)
        self.tape.append(bwd_apply)
        return o

    def _cross_attn_tc(self, q, kk, vv, nhead, rlen, p_drop):
        """cross_attn on the tcgen05 kernels (bf16 mode).  The per-head products become dense GEMMs against BLOCK-DIAGONAL token
        operands -- row h * Mp + m of the [heads * Mp, A] operand holds token m's head-h channels in head h's columns, zeros
        elsewhere -- so all heads share one launch of the generic tensor-core GEMM / weight-gradient kernels: 8x the useful
        flops at a tensor rate two orders above the CUDA cores', and the [frames x heads x tokens] probability tensors are bf16
        (half the HBM traffic, which is what bounds these launches).  The reductions over the frames (apply, dq) are the
        tcgen05 weight-gradient contraction; their block-diagonal part is the result."""
        B, M, A = q.v.shape
        slot = kk.v.shape[1]
        dh = A // nhead
        Mp = _round_up(M, 64 // math.gcd(nhead, 64))          # heads * Mp % 64 == 0: the K extent of a bf16 tensor-core source
        NP, bf = nhead * Mp, torch.bfloat16
        alpha = 1.0 / math.sqrt(dh)

        def blockdiag(x):                         # [B, M, A] -> [B, NP, A] bf16
            out = self.new((B, NP, A), bf, zero=True)
            out.view(B, nhead, Mp, nhead, dh).diagonal(dim1=1, dim2=3)[:, :M].copy_(x.view(B, M, nhead, dh).permute(0, 1, 3, 2))
            return out

        def diag_of(full):                        # [B, NP, A] -> the [B, M, dh, heads] view of its diagonal blocks
            return full.view(B, nhead, Mp, nhead, dh).diagonal(dim1=1, dim2=3)[:, :M]

        tcg = lambda A_, W_, N_, out, **kw: ops.gemm([S(A_, W_)], N_, out, len=rlen, tc=True, **kw)
        Qbd = blockdiag(q.v)
        L_ = self.new((B, slot, NP), bf)
        tcg(kk.v, Qbd, NP, L_, alpha=alpha)
        # zero beyond the valid rows: the tcgen05 weight-gradient contraction consumes whole 64-row stages
        Pv = Var(self.new((B, slot, NP), bf, zero=True), rlen)
        ops.col_softmax(L_, Pv.v, NP, len=rlen)
        del L_

        def bwd_softmax():
            if Pv.g is None:
                return
            dL = Pv.g                                   # in place
            ops.col_softmax_bwd(Pv.v, Pv.g, dL, NP, len=rlen)
            kg = self.G(kk)
            tcg(dL, dense(Qbd.transpose(1, 2)), A, kg, alpha=alpha, res=kg)
            full = self.new((B, NP, A))
            ops.wgrad(dL, kk.v, NP, A, full, len=rlen, alpha=alpha, accumulate=False, per_video=True, tc=True)
            self.G(q).view(B, M, nhead, dh).permute(0, 1, 3, 2).add_(diag_of(full))
        self.tape.append(bwd_softmax)
        Pd = self.dropout(Pv, p_drop)
        o = Var(self.new((B, M, A)))
        full = self.new((B, NP, A))
        ops.wgrad(Pd.v, vv.v, NP, A, full, len=rlen, accumulate=False, per_video=True, tc=True)
        o.v.view(B, M, nhead, dh).permute(0, 1, 3, 2).copy_(diag_of(full))
        del full

        def bwd_apply():
            if o.g is None:
                return
            OGbd = blockdiag(o.g)
            fresh = Pd.g is None
            if fresh:
                Pd.g = self.new(Pd.v.shape, bf, zero=True)
            tcg(vv.v, OGbd, NP, Pd.g, res=None if fresh else Pd.g)
            vg = self.G(vv)
            tcg(Pd.v, dense(OGbd.transpose(1, 2)), A, vg, res=vg)
        self.tape.append(bwd_apply)
        return o

    # ------------------------------------------------------------------ frame branch
    def _m2_fold(self, pfx, i, Lr, groups):
        """Folded MSTCN++ layer (see FactEngine._m2_fold) as a differentiable function of the six parameters."""
        def dense(name):
            w = self.P(name)
            if groups > 1:
                co, ci = w.shape[0] // groups, w.shape[1]
                full = w.new_zeros(w.shape[0], ci * groups, w.shape[2])
                for j in range(groups):
                    full[j * co:(j + 1) * co, j * ci:(j + 1) * ci] = w[j * co:(j + 1) * co]
                w = full
            return w.double()
        w1, w2 = dense(f'{pfx}conv_dilated_1.{i}.weight'), dense(f'{pfx}conv_dilated_2.{i}.weight')     # (F, F, 3)
        b1, b2 = self.P(f'{pfx}conv_dilated_1.{i}.bias').double(), self.P(f'{pfx}conv_dilated_2.{i}.bias').double()
        wf, bf = self.P(f'{pfx}conv_fusion.{i}.weight')[:, :, 0].double(), self.P(f'{pfx}conv_fusion.{i}.bias').double()
        F = w1.shape[0]
        A1, A2 = wf[:, :F], wf[:, F:]
        d1, d2 = 2 ** (Lr - 1 - i), 2 ** i
        acc = {}
        for k in range(3):
            acc[(k - 1) * d1] = acc.get((k - 1) * d1, 0) + A1 @ w1[:, :, k]
            acc[(k - 1) * d2] = acc.get((k - 1) * d2, 0) + A2 @ w2[:, :, k]
        Wt_ = torch.stack([acc[o] for o in self._m2_offsets(i, Lr)]).float()
        bt = (A1 @ b1 + A2 @ b2 + bf).float()
        return Wt_, bt

    def _m2_fold_all(self, pfx, Lr, groups):
        """Every layer of an MSTCN++ branch folded at once (the same algebra as _m2_fold, batched over the layers: a handful of
        batched fp64 products instead of six small ones per layer).  Needs five distinct tap offsets in every layer (even
        layer counts).  Returns ([Lr, 5, F, F] taps in _m2_offsets order, [Lr, F] biases)."""
        def dense_w(name):
            w = self.P(name)
            if groups > 1:
                co, ci = w.shape[0] // groups, w.shape[1]
                full = w.new_zeros(w.shape[0], ci * groups, w.shape[2])
                for j in range(groups):
                    full[j * co:(j + 1) * co, j * ci:(j + 1) * ci] = w[j * co:(j + 1) * co]
                w = full
            return w
        st = lambda fmt, f=self.P: torch.stack([f(pfx + fmt.format(i)) for i in range(Lr)]).double()
        w1, w2 = st('conv_dilated_1.{}.weight', dense_w), st('conv_dilated_2.{}.weight', dense_w)       # [L, F, F, 3]
        b1, b2, bf = st('conv_dilated_1.{}.bias'), st('conv_dilated_2.{}.bias'), st('conv_fusion.{}.bias')
        wf = st('conv_fusion.{}.weight')[:, :, :, 0]                                                         # [L, F, 2F]
        F = w1.shape[1]
        A1, A2 = wf[:, :, :F], wf[:, :, F:]
        P1 = torch.einsum('lof,lfik->lkoi', A1, w1)                                                          # [L, 3, F, F]
        P2 = torch.einsum('lof,lfik->lkoi', A2, w2)
        bt = torch.einsum('lof,lf->lo', A1, b1) + torch.einsum('lof,lf->lo', A2, b2) + bf
        c = P1[:, 1] + P2[:, 1]
        h = Lr // 2                  # layers i < h: dilation 2^(Lr-1-i) of conv_dilated_1 is the larger one, so its taps are the outer ones
        lo = torch.stack([P1[:h, 0], P2[:h, 0], c[:h], P2[:h, 2], P1[:h, 2]], dim=1)
        hi = torch.stack([P2[h:, 0], P1[h:, 0], c[h:], P1[h:, 2], P2[h:, 2]], dim=1)
        return torch.cat([lo, hi]).float(), bt.float()

    def frame_branch_t(self, pfx, bc, x, in_map, ln):
        F, H, Lr, C = bc['f_dim'], bc['hid_dim'], bc['f_layers'], self.ncls()
        m2, ng, p = bc['f'] == 'm2', bc['f_ngp'], float(bc['dropout'])
        cur = x
        if in_map:
            w = pfx + ('conv_1x1_in' if m2 else 'conv_1x1')
            cur = self.linear([src(x, self.taps_w(w + '.weight')[0])], F, bias=self.W(w + '.bias'), ln=ln, tag='in_proj')
        for i in range(Lr):
            if not m2:
                q = f'{pfx}layers.{i}.'
                w3, d = self.taps_w(q + 'conv_dilated.weight', ng), 2 ** i
                h = self.linear([src(cur, w3[k], off=(k - 1) * d) for k in range(3)], F, bias=self.W(q + 'conv_dilated.bias'),
                                relu=True, ln=ln, tag='tcn_conv3')
                w1 = self.taps_w(q + 'conv_1x1.weight')[0]
                if p > 0:
                    t = self.linear([src(h, w1)], F, bias=self.W(q + 'conv_1x1.bias'), ln=ln, tag='tcn_1x1')
                    nxt = self.add(cur, self.dropout(t, p))
                else:
                    nxt = self.linear([src(h, w1)], F, bias=self.W(q + 'conv_1x1.bias'), res=cur, ln=ln, tag='tcn_1x1')
                if bc['f_ln']:
                    nxt = self.layernorm(nxt, q + 'norm.weight', q + 'norm.bias', ln=ln)
            else:
                offs = self._m2_offsets(i, Lr)
                if Lr % 2 == 0:
                    Wall, ball = self.Dn(('m2fold_all', pfx), lambda: self._m2_fold_all(pfx, Lr, ng))
                    Wf, bf = Wall[i], ball[i]
                else:
                    Wf, bf = self.Dn(('m2fold', pfx, i), lambda: self._m2_fold(pfx, i, Lr, ng))
                f = self.linear([src(cur, Wf[j], off=o) for j, o in enumerate(offs)], F, bias=bf, relu=True, ln=ln, tag='tcn_m2')
                if i != Lr - 1:
                    f = self.dropout(f, p)
                nxt = self.add(f, cur)
            cur = nxt
        out = self.linear([src(cur, self.taps_w(pfx + 'conv_out.weight')[0])], H, bias=self.W(pfx + 'conv_out.bias'), ln=ln, tag='conv_out')
        return self.splice(out, C, ln=ln, want_pred=True)

    # ------------------------------------------------------------------ token side
    def _mha_block(self, pfx, x, qpos, nhead, p):
        """q = k = x + pos, v = x through the packed in_proj (basic.py:437,500) + out_proj; returns the attention output."""
        A = x.v.shape[-1]
        Win, bin_ = self.W(pfx + 'in_proj_weight'), self.W(pfx + 'in_proj_bias')
        xq = self.addpos(x, qpos)
        qk = self.linear([src(xq, Win[:2 * A])], 2 * A, bias=bin_[:2 * A])
        v = self.linear([src(x, Win[2 * A:])], A, bias=bin_[2 * A:])
        o = self.mha_self(self.cols(qk, 0, A), self.cols(qk, A, 2 * A), v, nhead, p)
        return self.linear([src(o, self.W(pfx + 'out_proj.weight'))], A, bias=self.W(pfx + 'out_proj.bias'))

    def _ffn(self, q, x, p):
        ff = self.P(q + 'linear1.weight').shape[0]
        A = x.v.shape[-1]
        h = self.linear([src(x, self.W(q + 'linear1.weight'))], ff, bias=self.W(q + 'linear1.bias'), relu=True)
        h = self.dropout(h, p)
        return self.linear([src(h, self.W(q + 'linear2.weight'))], A, bias=self.W(q + 'linear2.bias'))

    def sca_decoder_t(self, pfx, bc, frame, rlen, pos_idx=None):
        B, M, A, H, nh, p = self.B, self.ntok, bc['a_dim'], bc['hid_dim'], bc['a_nhead'], float(bc['dropout'])
        qpos = self.qpos_w()
        tgt = Var(self.new((B, M, A), zero=True), needs_grad=False)
        for i in range(bc['a_layers']):
            q = f'{pfx}layers.{i}.'
            t2 = self._mha_block(q + 'self_attn.', tgt, qpos, nh, p)
            tgt = self.layernorm(self.dropout(t2, p), q + 'norm1.weight', q + 'norm1.bias', res=tgt)
            c = q + 'multihead_attn.'
            cb = self.W(c + 'in_proj_bias')
            if (c + 'in_proj_weight') in self._params:
                Wc = self.W(c + 'in_proj_weight')
                wq, wk, wv = Wc[:A], Wc[A:2 * A], Wc[2 * A:]
            else:
                wq, wk, wv = self.W(c + 'q_proj_weight'), self.W(c + 'k_proj_weight'), self.W(c + 'v_proj_weight')
            cq = self.linear([src(self.addpos(tgt, qpos), wq)], A, bias=cb[:A])
            # keys / values in the activation dtype (bf16 in bf16 mode, like the inference engine): their projections, data and
            # weight gradients then run on the tensor cores
            kk = self.linear([src(self.with_pos_t(frame, rlen, pos_idx), wk)], A, bias=cb[A:2 * A], ln=rlen, tag='sca_kv')
            vv = self.linear([src(frame, wv)], A, bias=cb[2 * A:], ln=rlen, tag='sca_kv')
            o = self.cross_attn(cq, kk, vv, nh, rlen, p)
            t2 = self.linear([src(o, self.W(c + 'out_proj.weight'))], A, bias=self.W(c + 'out_proj.bias'))
            tgt = self.layernorm(self.dropout(t2, p), q + 'norm2.weight', q + 'norm2.bias', res=tgt)
            t2 = self._ffn(q, tgt, p)
            tgt = self.layernorm(self.dropout(t2, p), q + 'norm3.weight', q + 'norm3.bias', res=tgt)
        t = self.layernorm(tgt, pfx + 'norm.weight', pfx + 'norm.bias')
        return self.linear([src(t, self.W(pfx + 'out_linear.weight'))], H, bias=self.W(pfx + 'out_linear.bias'))

    def sa_decoder_t(self, pfx, bc, x):
        A, H, nh, p = bc['a_dim'], bc['hid_dim'], bc['a_nhead'], float(bc['dropout'])
        qpos = self.qpos_w()
        for i in range(bc['a_layers']):
            q = f'{pfx}layers.{i}.'
            t2 = self._mha_block(q + 'multihead_attn.', x, qpos, nh, p)
            x = self.layernorm(self.dropout(t2, p), q + 'norm1.weight', q + 'norm1.bias', res=x)
            t2 = self._ffn(q, x, p)
            x = self.layernorm(self.dropout(t2, p), q + 'norm2.weight', q + 'norm2.bias', res=x)
        return self.linear([src(x, self.W(pfx + 'out_linear.weight'))], H, bias=self.W(pfx + 'out_linear.bias'))

    def qpos_w(self):
        h = self._wt.get('qpos')
        if h is None:
            a = self.W('action_query')
            h = self._wt['qpos'] = Wt(a.w[:, 0], a.g[:, 0])
        return h

    # ------------------------------------------------------------------ X2Y_map (basic.py:349-389)
    def _x2y_logit(self, pfx, rows, rlen, pos_idx, action, q_name, k_name):
        """logit[t, m] = (rows[t] + pos) . (alpha Wr^T tq[m]) + alpha tq[m] . br with tq = Linear_t(action + qpos): the
        projection of all rows (Linear_r, weight Wr / bias br) is never formed.  f2a: t = Y_Q, r = X_K; a2f: t = X_K, r = Y_Q."""
        H, M = rows.v.shape[-1], self.ntok
        alpha = 1.0 / math.sqrt(H)
        tq = self.linear([src(self.addpos(action, self.qpos_w()), self.W(pfx + q_name + '.weight'))], H, bias=self.W(pfx + q_name + '.bias'))
        WrT = self.D(('tr', pfx + k_name), lambda: self.P(pfx + k_name + '.weight').t())
        br = self.D(('b1', pfx + k_name), lambda: self.P(pfx + k_name + '.bias')[None, :])
        qt = self.linear([src(tq, WrT)], H, alpha=alpha)
        cb = self.linear([src(tq, br)], 1, alpha=alpha)
        Mp = _round_up(M, 4)
        logit = Var(self.new((self.B, self.slot, Mp), zero=True), rlen)
        cbv = Var(cb.v[:, :, 0], None, None)
        rp = self.with_pos_t(rows, rlen, pos_idx)
        tc = self.mode == 'bf16' and self.use_tc and H % 64 == 0           # per-video operand: bf16 or tf32 tcgen05 GEMM
        Wq = qt.v if (not tc or rp.v.dtype == torch.float32) else qt.v.to(rp.v.dtype)
        ops.gemm([S(rp.v, Wq)], M, logit.v, len=rlen, bias=cbv.v, tc=tc)

        def bwd():
            if logit.g is None:
                return
            ops.colsum(logit.g, M, self.G(cb)[:, :, 0], len=rlen, per_video=True)
            ops.wgrad(logit.g, rp.v, M, H, self.G(qt), len=rlen, per_video=True)
            if rp.needs_grad:
                qT = self.new((self.B, H, Mp), zero=True)
                ops.transpose(qt.v, qT[:, :, :M])
                rg = self.G(rp)
                self.mm([S(logit.g, qT, K=M)], H, rg, len=rlen, res=rg)
        self.tape.append(bwd)
        return logit

    def f2a_t(self, pfx, bc, rows, rlen, pos_idx, action):
        M, A, H, p = self.ntok, bc['a_dim'], bc['hid_dim'], float(bc['dropout'])
        logit = self._x2y_logit(pfx, rows, rlen, pos_idx, action, 'Y_Q', 'X_K')
        Mp = logit.v.shape[-1]
        attn = Var(self.new((self.B, self.slot, Mp), zero=True), rlen)
        ops.col_softmax(logit.v, attn.v, M, len=rlen)
        xbar = Var(self.new((self.B, M, H)))
        ops.wgrad(attn.v, rows.v, M, H, xbar.v, len=rlen, accumulate=False, per_video=True)

        def bwd():
            if xbar.g is None:
                return
            dA = self.new((self.B, self.slot, Mp), zero=True)
            ops.gemm([S(rows.v, xbar.g)], M, dA, len=rlen)
            if rows.needs_grad:
                xT = self.new((self.B, H, Mp), zero=True)
                ops.transpose(xbar.g, xT[:, :, :M])
                rg = self.G(rows)
                self.mm([S(attn.v, xT, K=M)], H, rg, len=rlen, res=rg)
            ops.col_softmax_bwd(attn.v, dA, self.G(logit), M, len=rlen, accumulate=True)
        self.tape.append(bwd)
        feat = self.linear([src(xbar, self.W(pfx + 'X_V.weight'))], H, bias=self.W(pfx + 'X_V.bias'))
        Wy = self.W(pfx + 'Y_W.weight')
        out = self.linear([src(self.dropout(action, p), Wy[:, :H]), src(self.dropout(feat, p), Wy[:, H:])], A, bias=self.W(pfx + 'Y_W.bias'))
        return out, logit, attn

    def a2f_t(self, pfx, bc, action, rows, rlen, pos_idx):
        M, F, H, p = self.ntok, bc['f_dim'], bc['hid_dim'], float(bc['dropout'])
        logit = self._x2y_logit(pfx, rows, rlen, pos_idx, action, 'X_K', 'Y_Q')
        Mp = logit.v.shape[-1]
        attn = Var(self.new((self.B, self.slot, Mp), zero=True), rlen)
        ops.row_softmax(logit.v, attn.v, M, len=rlen)

        def bwd():
            if attn.g is None:
                return
            ops.row_softmax_bwd(attn.v, attn.g, self.G(logit), M, len=rlen, accumulate=True)
        self.tape.append(bwd)
        xv = self.linear([src(action, self.W(pfx + 'X_V.weight'))], H, bias=self.W(pfx + 'X_V.bias'))
        xvT = Var(self.new((self.B, H, Mp), zero=True))
        ops.transpose(xv.v, xvT.v[:, :, :M])

        def bwd_t():
            if xvT.g is None:
                return
            t = self.new((self.B, M, H))
            ops.transpose(xvT.g[:, :, :M], t)
            ops.ew(ops.EW_AXPY, t, self.G(xv), H)
        self.tape.append(bwd_t)
        u = self.linear([src(attn, xvT, K=M)], H, ln=rlen, dtype=torch.float32)
        Wy = self.W(pfx + 'Y_W.weight')
        out = self.linear([src(self.dropout(rows, p), Wy[:, :H]), src(self.dropout(u, p), Wy[:, H:])], F, bias=self.W(pfx + 'Y_W.bias'),
                          ln=rlen, tag='x2y_rows')
        return out, logit, attn

    # ------------------------------------------------------------------ temporal down / up-sampling
    def downsample_t(self, i, bc, frame, pred, st):
        pfx = f'block_list.{i}.'
        B, slot, H = self.B, self.slot, bc['hid_dim']
        Hh, I32 = H // 2, torch.int32
        seg = dict(label=self.new((B, slot), I32), start=self.new((B, slot), I32), len=self.new((B, slot), I32),
                   center=self.new((B, slot), I32), nseg=self.new((B,), I32))
        ops.tdu_segment(pred, seg['label'], seg['start'], seg['len'], seg['center'], seg['nseg'], len=self.len)
        nseg = seg['nseg']
        st.update(seg_label=seg['label'], seg_lens=seg['len'], seg_start=seg['start'], nseg=nseg, tdu_pred=pred)
        seg0 = Var(self.new((B, slot, H), frame.v.dtype), nseg)
        ops.segment_mean(frame.v, seg0.v, seg['label'], seg['start'], seg['len'], nseg)

        def bwd_mean():
            if seg0.g is None:
                return
            ops.segment_expand(seg0.g, seg['label'], seg['len'], self.G(frame), H, len=self.len, inv_len=True, accumulate=True)
        self.tape.append(bwd_mean)
        g = pfx + 'seg_update.'
        cur = seg0
        for l in range(self.hp['s_layers']):
            Wih = self.D(('cat', g, 'w_ih', l), lambda l=l: torch.cat([self.P(f'{g}weight_ih_l{l}'), self.P(f'{g}weight_ih_l{l}_reverse')], 0))
            bih = self.D(('cat', g, 'b_ih', l), lambda l=l: torch.cat([self.P(f'{g}bias_ih_l{l}'), self.P(f'{g}bias_ih_l{l}_reverse')], 0))
            gi = self.linear([src(cur, Wih)], 6 * Hh, bias=bih, ln=nseg, dtype=torch.float32, tag='gru_in')
            cur = self.gru(g, l, gi, nseg)
        hr = self.relu(cur)
        seg2 = self.linear([src(hr, self.W(pfx + 'seg_combine.weight'))], H, bias=self.W(pfx + 'seg_combine.bias'), ln=nseg, tag='seg_combine')
        segf, st['seg_clogit_v'], _ = self.splice(seg2, self.ncls(), ln=nseg)
        return segf, seg

    def gru(self, g, l, gi, nseg):
        """One bidirectional GRU layer (forward: the fp32 cluster kernel; backward: gate pre-activations of all steps by one
        GEMM over the saved hidden states, then BPTT, then weight / bias gradients by the generic kernels)."""
        B, slot = gi.v.shape[0], gi.v.shape[1]
        Whf, Whb = self.W(f'{g}weight_hh_l{l}'), self.W(f'{g}weight_hh_l{l}_reverse')
        bhf, bhb = self.W(f'{g}bias_hh_l{l}'), self.W(f'{g}bias_hh_l{l}_reverse')
        Hh = Whf.w.shape[1]
        y = Var(self.new((B, slot, 2 * Hh), zero=True), nseg)
        # bf16 mode: the inference engine's tensor-core recurrence (bf16 weights and exchanged state, flag-in-data exchange: about
        # half the per-step latency of the fp32 cluster kernel); the backward pass re-derives the gates from the saved states
        ops.gru_bidir(gi.v, Whf.w, bhf.w, Whb.w, bhb.w, y.v, nseg, relu=False, mma=(self.mode == 'bf16' and self.use_tc and Hh == 256))

        def bwd():
            if y.g is None:
                return
            gh = self.new((B, slot, 6 * Hh))
            ops.gemm([S(y.v[:, :, :Hh], Whf.w, off=-1)], 3 * Hh, gh[:, :, :3 * Hh], len=nseg, bias=bhf.w)
            ops.gemm([S(y.v[:, :, Hh:], Whb.w, off=1)], 3 * Hh, gh[:, :, 3 * Hh:], len=nseg, bias=bhb.w)
            gi.g = self.new((B, slot, 6 * Hh), zero=True)
            dgh = self.new((B, slot, 6 * Hh), zero=True)
            ops.gru_bwd(gi.v, gh, y.v, y.g, Whf.w, Whb.w, gi.g, dgh, nseg)
            def param_grads():
                self.wgrad(dgh[:, :, :3 * Hh], y.v[:, :, :Hh], 3 * Hh, Hh, Whf.g, off=-1, len=nseg)
                self.wgrad(dgh[:, :, 3 * Hh:], y.v[:, :, Hh:], 3 * Hh, Hh, Whb.g, off=1, len=nseg)
                ops.colsum(dgh[:, :, :3 * Hh], 3 * Hh, bhf.g, len=nseg)
                ops.colsum(dgh[:, :, 3 * Hh:], 3 * Hh, bhb.g, len=nseg)
            self.pgrad(param_grads, dgh, y.v)
        self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------ blocks
    def token_splice_t(self, action):
        y, cl, _ = self.splice(action, self.ncls(tokens=True))
        return y, cl

    def input_block_t(self, i, bc, x, st):
        pfx = f'block_list.{i}.'
        frame, st['frame_clogit_v'], st['pred'] = self.frame_branch_t(pfx + 'frame_branch.', bc, x, True, self.len)
        action = self.sca_decoder_t(pfx + 'action_branch.', bc, frame, self.len)
        action, st['action_clogit_v'] = self.token_splice_t(action)
        return frame, action

    def update_block_t(self, i, bc, frame, action, st):
        pfx = f'block_list.{i}.'
        tok, st['f2a_logit_v'], f2a_attn = self.f2a_t(pfx + 'f2a_layer.', bc, frame, self.len, None, action)
        st['f2a_attn'] = f2a_attn.v
        action = self.sa_decoder_t(pfx + 'action_branch.', bc, tok)
        action, st['action_clogit_v'] = self.token_splice_t(action)
        fr, st['a2f_logit_v'], a2f_attn = self.a2f_t(pfx + 'a2f_layer.', bc, action, frame, self.len, None)
        st['a2f_attn'] = a2f_attn.v
        frame, st['frame_clogit_v'], st['pred'] = self.frame_branch_t(pfx + 'frame_branch.', bc, fr, False, self.len)
        return frame, action

    def update_block_tdu_t(self, i, bc, frame, action, pred, st):
        pfx = f'block_list.{i}.'
        F = bc['f_dim']
        seg2, seg = self.downsample_t(i, bc, frame, pred, st)
        nseg = seg['nseg']
        pidx = seg['center'] if self.frame_pos is not None else None
        tok, st['f2a_logit_v'], f2a_attn = self.f2a_t(pfx + 'f2a_layer.', bc, seg2, nseg, pidx, action)
        st['f2a_attn_seg'] = f2a_attn.v
        action = self.sa_decoder_t(pfx + 'action_branch.', bc, tok)
        action, st['action_clogit_v'] = self.token_splice_t(action)
        seg3, st['a2f_logit_v'], a2f_attn = self.a2f_t(pfx + 'a2f_layer.', bc, action, seg2, nseg, pidx)
        st['a2f_attn_seg'] = a2f_attn.v
        Wm = self.W(pfx + 'sf_merge.0.weight')                                  # [F, F+H], input = cat[s2f, frame]
        s2f = self.linear([src(seg3, Wm[:, :F])], F, ln=nseg, dtype=torch.float32, tag='sf_merge')
        fr = self.linear([src(frame, Wm[:, F:])], F, bias=self.W(pfx + 'sf_merge.0.bias'), relu=True, pre=s2f, pre_seg=seg, ln=self.len,
                         tag='sf_merge')
        frame, st['frame_clogit_v'], st['pred'] = self.frame_branch_t(pfx + 'frame_branch.', bc, fr, False, self.len)
        return frame, action

    # ------------------------------------------------------------------ whole step
    @staticmethod
    def time_mask_spans(T, cfg_tm):
        """basic.time_mask (basic.py:10-36) with replace_with_zero=True: TM.m spans, each of length
        min(int(TM.p * T), randrange(TM.t)) at start randrange(T - length); a zero-length draw ENDS the masking (the
        reference returns there).  Same calls to Python's ``random`` in the same order as the reference."""
        spans = []
        for _ in range(int(cfg_tm.m)):
            t = random.randrange(0, int(cfg_tm.t))
            t = min(int(float(cfg_tm.p) * T), t)
            t0 = random.randrange(0, T - t)
            if t == 0:
                break
            spans.append((t0, t0 + t))
        return spans

    def forward_train(self, seqs, forced_preds=None):
        """Train-mode forward of a batch.  Returns the ``out`` dict of FactEngine.run_packed (values) with the Vars of every
        tensor the loss reads under ``*_v`` keys.

        ``use_graphs``: the step is captured ONCE per batch shape (B, slot) into CUDA graphs -- one for the forward pass, one
        per section of the backward pass (the gradient-ready hooks run between them) -- and replayed afterwards: the ~4000
        kernel launches of a step cost a handful of graph launches instead of ~40 us of host work each.  Everything that
        changes from step to step enters through static device buffers: the features, the lengths, the dropout seed (the
        kernels add a device scalar to their seed), the time mask (one factor per frame), the loss gradients."""
        self._refresh_weights()
        hp, cfg = self.hp, self.m.cfg
        self.ntok, self.action_init, self.transcript = hp['ntoken'], None, None
        lengths = [int(s.shape[0]) for s in seqs]
        B, slot, D = len(seqs), _round_up(max(lengths), 128), hp['in_dim']
        self._set_arena(('train', B, slot))
        self.B, self.slot, self.keep = B, slot, True
        st = self._steps.get((B, slot))
        if st is None:
            if len(self._steps) >= 2:                       # every entry pins a whole step's activations
                self._steps.pop(next(iter(self._steps)))
            st = self._steps[(B, slot)] = dict(
                x=torch.zeros((B, slot, D), dtype=torch.float32, device=self.dev), ln=torch.zeros((B,), dtype=torch.int32, device=self.dev),
                tm=torch.ones((B, slot), dtype=torch.float32, device=self.dev) if cfg.TM.use else None, lengths=None, warm=False,
                ln_host=torch.zeros((B,), dtype=torch.int32).pin_memory())
        # ---- staging (never captured): features, lengths, seed, time mask into the static buffers
        if st['lengths'] != lengths and st['lengths'] is not None:
            st['x'].zero_()
        for b_, s_ in enumerate(seqs):
            st['x'][b_, :lengths[b_]].copy_(s_, non_blocking=True)
        st['ln_host'].copy_(torch.tensor(lengths, dtype=torch.int32))
        st['ln'].copy_(st['ln_host'], non_blocking=True)
        st['lengths'] = lengths
        self.step_no += 1
        self.seed_dev.fill_(self.step_no)
        if st['tm'] is not None:                          # basic.time_mask spans as one keep factor per frame
            tm = torch.ones((B, slot), dtype=torch.float32)
            for b_, T in enumerate(lengths):
                for t0, t1 in self.time_mask_spans(T, cfg.TM):
                    tm[b_, t0:t1] = 0.0
            st['tm'].copy_(tm, non_blocking=True)
        self.len, self._cur = st['ln'], st
        graphable = self.use_graphs and forced_preds is None
        if graphable and st.get('gF') is not None:
            if any(p_.grad is not None and p_.grad.data_ptr() == self._pg[n].data_ptr() for n, p_ in self._params.items()):
                raise RuntimeError('graph-captured training step: call optimizer.zero_grad(set_to_none=True) between steps (the '
                                   'gradient buffers are static)')
            for f in self._flat:
                f.zero_()
            st['gF'].replay()
            st['out']['lengths'] = lengths
            # backward: replay its graphs when they exist, else capture them from the closures of the forward capture
            self.tape = st['tape_done'] if st.get('gB') is not None else st['tape']
            return st['out']
        self.begin()
        if graphable and st['warm']:
            if self._cap_stream is None:
                self._cap_stream = torch.cuda.Stream(device=self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._cap_stream):
                out = self._forward_body(st, lengths, None)
            st['gF'], st['out'], st['tape'] = g, out, self.tape
            g.replay()
        else:
            out = self._forward_body(st, lengths, forced_preds)
            st['warm'] = True
        return out

    def _forward_body(self, st, lengths, forced_preds):
        hp, cfg = self.hp, self.m.cfg
        B, slot = self.B, self.slot
        ln = st['ln']
        C, H = hp['n_classes'], hp['blocks'][0]['hid_dim']
        self.frame_pos = self.const_table(('pe', slot, H), lambda: _pos_table(H, slot, self.dev)) if hp['fpos'] else None
        xin = Var(st['x'], ln, needs_grad=False)
        # nn.Dropout2d over whole feature channels (blocks.py:614-617); in bf16 mode the same pass casts the features to bf16
        xin = self.dropout(xin, float(cfg.FACT.cmr), channel=True, dtype=self.act)
        if st['tm'] is not None:                           # time_mask on the (already channel-masked) features (blocks.py:619-622)
            y = Var(torch.zeros_like(xin.v), ln, needs_grad=False)
            ops.ew(ops.EW_ROWSCALE, xin.v, y.v, xin.v.shape[-1], r=st['tm'].unsqueeze(-1), len=ln)
            xin = y
        frame, action, stash, u, pred = xin, None, [], 0, None
        for i, bc in enumerate(hp['blocks']):
            st = {}
            self.mark_section(i)
            if bc['type'] == 'i':
                frame, action = self.input_block_t(i, bc, frame, st)
            elif bc['type'] == 'u':
                frame, action = self.update_block_t(i, bc, frame, action, st)
            elif bc['type'] == 'U':
                if forced_preds is not None:
                    fp = self.new((B, slot), torch.int32, zero=True)
                    for b in range(B):
                        fp[b, :lengths[b]].copy_(forced_preds[u][b].to(torch.int32), non_blocking=True)
                    pred = fp
                frame, action = self.update_block_tdu_t(i, bc, frame, action, pred, st)
                u += 1
            else:
                raise NotImplementedError(f'training step: block type {bc["type"]!r}')
            pred = st['pred']
            st['frame_feature'], st['action_feature'] = frame.v, action.v
            for k in ('frame_clogit', 'action_clogit', 'seg_clogit', 'f2a_logit', 'a2f_logit'):
                if k + '_v' in st:      # plain tensors under the inference engine's keys (LossRunner, stash_video, eval)
                    st[{'f2a_logit': 'f2a_attn_logit', 'a2f_logit': 'a2f_attn_logit'}.get(k, k)] = st[k + '_v'].v
            stash.append(st)
        out = dict(blocks=stash, lengths=lengths)
        last = stash[-1]
        if self.clip and self.m.text_embeddings is not None:
            self.mark_section(len(hp['blocks']))
            P = self.P('frame_projection.projection.0.weight').shape[0]
            w0 = self.D(('clip_w0pad',), lambda: torch.nn.functional.pad(self.P('frame_projection.projection.0.weight'), (0, C)))
            h1 = self.linear([src(frame, w0)], P, bias=self.W('frame_projection.projection.0.bias'), ln=ln, tag='clip')
            h2 = self.layernorm(h1, 'frame_projection.projection.1.weight', 'frame_projection.projection.1.bias', relu=True, ln=ln)
            h2 = self.dropout(h2, float(cfg.CLIP.projection_dropout))
            e0 = self.linear([src(h2, self.W('frame_projection.projection.4.weight'))], 512, bias=self.W('frame_projection.projection.4.bias'),
                             ln=ln, dtype=torch.float32, tag='clip')
            emb = self.l2norm(e0, ln=ln)
            sim = self.linear([src(emb, self.const(self.m.text_embeddings.detach()))], C, alpha=1.0 / hp['temp'], ln=ln, dtype=torch.float32,
                              tag='clip')
            out['projected_frame_embeddings'], out['clip_logit'], out['clip_logit_v'] = emb.v, sim.v, sim
            flogit = sim.v
        else:
            flogit = last['frame_clogit']
        pred64 = self.new((B, slot), torch.int64, zero=True)
        M = self.ntok
        if 'a2f_attn' in last:
            ops.fuse_eval(last['action_clogit'], last['a2f_attn'], flogit, hp['mwt'], pred64, M, C, len=ln)
        elif 'a2f_attn_seg' in last:
            ops.fuse_eval(last['action_clogit'], last['a2f_attn_seg'], flogit, hp['mwt'], pred64, M, C, seg_label=last['seg_label'], len=ln)
        else:
            ops.fuse_eval(None, None, flogit, hp['mwt'], pred64, 0, C, len=ln)
        out['pred'] = pred64
        return out

    def backward(self, seeds):
        """seeds: list of (Var, gradient tensor).  Runs the tape; returns {parameter name: gradient} (fp32)."""
        st = self._cur
        if st is not None and st.get('gB') is not None and self.tape is st.get('tape_done'):       # replay
            for buf_, (_, g) in zip(st['seed_bufs'], seeds):
                buf_.copy_(g)
            for g, k in st['gB']:
                g.replay()
                if k is not None:
                    self._section_done(k)
            return self._pg
        capture = st is not None and st.get('gF') is not None and st.get('tape') is self.tape and self.use_graphs
        if capture:                                         # loss gradients enter through static buffers
            st['seed_bufs'] = [torch.zeros_like(g) for _, g in seeds]
            for buf_, (var, g) in zip(st['seed_bufs'], seeds):
                buf_.copy_(g)
                assert var.g is None
                var.g = buf_
            st['gB'] = []
        else:
            for var, g in seeds:
                if var.g is None:
                    var.g = g
                else:
                    var.g += g
        # segments of the reversed tape, each ending at a section marker
        segs, cur = [], []
        for item in reversed(self.tape):
            if isinstance(item, tuple):
                segs.append((cur, item[1], item[2]))
                cur = []
            else:
                cur.append(item)
        if cur:                                             # entries before the first section (input augmentations): no parameters
            segs.append((cur, None, []))
        for fns, k, derived in segs:
            if capture:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._cap_stream):
                    for fn in fns:
                        fn()
                    self._side_join()
                    self._finish_derived(derived)
                g.replay()
                st['gB'].append((g, k))
            else:
                for fn in fns:
                    fn()
                self._side_join()
                self._finish_derived(derived)
            if k is not None:
                self._section_done(k)
        if capture:
            st['tape_done'] = self.tape = object()          # the closures live on in the graphs; later steps replay
            st['tape'] = None
        else:
            self.tape = []
        self._derived, self._wt, self._leaf = [], {}, {}
        return self._pg
