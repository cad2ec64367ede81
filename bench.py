"""Benchmark of the FACT_CLIP forward hot path (BASELINE.json metric: frames/sec, T=4096, 2048-d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--videos B] [--mode bf16|fp32]

One "step" = one batched forward of B synthetic videos of T=4096 frames per GPU through the
havid_view0_lh_pt_holdout FACT_CLIP configuration (random-init weights, segment-structured synthetic
features).  Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_FRAMES, IN_DIM, N_CLASSES = 4096, 2048, 75
PRESET = 'havid_view0_lh_pt_holdout'
METRIC, UNIT = 'frames/sec FACT_CLIP fwd (T=4096,2048-d)', 'frames/s'


# dram__bytes_read.sum + dram__bytes_write.sum of one tcn_layer_kernel launch (ncu --set full), keyed by videos per step;
# algorithmic bytes at 64 videos: x in + y out = 2 * 64 * 4096 * 256 * 2 B = 268.4 MB
TCN_DRAM_BYTES_PER_LAUNCH = {64: 227.36e6}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        mx = max((float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()), default=None)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=reasons, samples=len(sm))


def build_model(mode):
    from fact_clip_b200 import config as C
    from fact_clip_b200.models.blocks import FACT_CLIP
    from fact_clip_b200.utils.synth import make_text_embeddings
    cfg = C.PRESETS[PRESET]()
    torch.manual_seed(0)
    net = FACT_CLIP(cfg, IN_DIM, N_CLASSES, make_text_embeddings(N_CLASSES)).eval()
    net.compute_mode = mode
    return net, cfg


def alg_flops_tcn_layer(F):
    """Dilated residual layer: conv3 (3 taps) + 1x1 = 8 F^2 FLOP per frame."""
    return 8.0 * F * F


def stock_reference():
    """The UNMODIFIED reference model (baseline/_ref, installed by tools/install_reference.py) on the CPU, or None when it is
    not installed.  yacs is not in the image: oracle/_yacs_shim stands in for it (the one other thing bench.py takes from
    oracle/).  Same constructor call as scripts/run_eval.py:124-125 with the preset's values in the reference's own cfg."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref, 'fact_clip')):
        return None
    try:
        import contextlib
        import io
        import warnings
        sys.path.insert(0, os.path.join(ROOT, 'oracle', '_yacs_shim'))
        sys.path.insert(0, ref)
        warnings.filterwarnings('ignore')
        from fact_clip.configs.default import get_cfg_defaults
        from fact_clip.models.blocks import FACT_CLIP as RefFactClip
        from fact_clip_b200 import config as C
        from fact_clip_b200.utils.synth import make_text_embeddings
        ours, cfg = C.PRESETS[PRESET](), get_cfg_defaults()
        for sect in ('FACT', 'Bi', 'Bu', 'BU', 'CLIP', 'TM'):
            for k, v in ours[sect].items():
                cfg[sect][k] = v
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            net = RefFactClip(cfg, IN_DIM, N_CLASSES, make_text_embeddings(N_CLASSES)).eval()
        return net
    except Exception as e:          # keep the bench alive: fall back to the port and say why
        sys.stderr.write(f'stock reference unavailable ({type(e).__name__}: {e}); timing the oracle port instead\n')
        return None


def cpu_baseline(n_videos, threads, state={}):
    """The reference's CPU path timed on the host cores: the stock reference model when baseline/_ref is installed
    ('kind: reference'), else the oracle port (oracle/fact_oracle.py, 'kind: port').  -> (frames/s, seconds, kind)."""
    from fact_clip_b200.utils.synth import make_batch
    torch.set_num_threads(threads)
    xs, ys = make_batch([T_FRAMES] * n_videos, IN_DIM, N_CLASSES, base_seed=1000)
    if 'net' not in state:
        state['net'] = stock_reference()
    ref = state['net']
    if ref is not None:
        import contextlib
        import io
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            if 'warm' not in state:
                ref(xs[:1], ys[:1])
                state['warm'] = True
            t0 = time.perf_counter()
            ref(xs, ys)
            dt = time.perf_counter() - t0
        return n_videos * T_FRAMES / dt, dt, 'reference'
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import fact_oracle as O
    if 'sd' not in state:
        net, cfg = build_model('fp32')
        state['sd'] = {k: v.detach().clone() for k, v in net.state_dict().items()}
        state['hp'] = O.hparams_from_cfg(cfg, IN_DIM, N_CLASSES)
        O.forward(state['sd'], state['hp'], xs[:1], clip=True)                       # warm-up
    t0 = time.perf_counter()
    O.forward(state['sd'], state['hp'], xs, clip=True)
    dt = time.perf_counter() - t0
    return n_videos * T_FRAMES / dt, dt, 'port'


def parity_vs_reference(net, n_videos=2):
    """Free-running bf16 parity of THIS run against the stock reference (same leg as the CPU baseline: the only place the bench
    touches baseline/_ref): per-block logits (relative L2, blocks before the first segmentation divergence), |dS| per U block
    and the final per-frame argmax agreement on the first videos of the CPU sample.  None when the reference is not installed."""
    import contextlib
    import io
    ref = stock_reference()
    if ref is None:
        return None
    from fact_clip_b200.utils.synth import make_batch
    xs, ys = make_batch([T_FRAMES] * n_videos, IN_DIM, N_CLASSES, base_seed=1000)
    dev = next(net.parameters()).device
    keep = getattr(net, 'keep_attn', False)
    net.keep_attn = True
    ours = net([x.to(dev) for x in xs], [y.to(dev) for y in ys])
    worst, ds, agree, tot = 0.0, [], 0, 0
    rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm().clamp_min(1e-12))
    for b, (x, y) in enumerate(zip(xs, ys)):
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            r = ref([x], [y])
        net.stash_video(b)
        same = True
        for blk, rb in zip(net.block_list, ref.block_list):
            if hasattr(rb, 'tdu'):
                ds.append(abs(int(rb.tdu.num_seg) - int(blk.tdu.num_seg)))
                same = same and torch.equal(blk.tdu.seg_label.cpu(), rb.tdu.seg_label.cpu())
            if same:
                worst = max(worst, rel(blk.frame_clogit, rb.frame_clogit), rel(blk.action_clogit, rb.action_clogit))
        agree += int((ours[b]['pred'] == r[0]['pred']).sum())
        tot += len(r[0]['pred'])
    net.keep_attn = keep
    return {'videos': n_videos, 'mode': 'free-running (no teacher forcing)', 'worst_logit_rel_l2_before_divergence': worst,
            'abs_dS_per_U_block': ds, 'argmax_agreement': agree / tot}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 4               # videos per step: a bounded sample of the workload (a few seconds of CPU work per step on 16 cores)
    vals = []
    for _ in range(args.warmup + args.steps):
        v, dt, kind = cpu_baseline(n, threads)
        vals.append((v, dt))
    vals = vals[args.warmup:]
    fps = sum(n * T_FRAMES for _ in vals) / sum(dt for _, dt in vals)
    what = ('stock reference FACT_CLIP (baseline/_ref, eval, no_grad, torch CPU fp32)' if kind == 'reference'
            else 'oracle port (torch CPU fp32)')
    sample = f'{n} videos x T={T_FRAMES} per step, {what}, {threads} threads'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * sum(dt for _, dt in vals) / len(vals), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'FACT_CLIP {PRESET} forward, T={T_FRAMES}, D={IN_DIM}, C={N_CLASSES}, {n} videos/step (CPU)'},
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def calibrated_regime(net, cfg, dev, x, ln, lengths, B, T, args):
    """Second, clearly labelled regime: the SAME forward workload with weights after a short run of this repo's own training step
    on synthetic videos of the same distribution.  Random-init weights make the TDU segmentation pathological (hundreds to
    thousands of one-frame segments per video), which inflates the latency-bound GRU chain; a few dozen optimiser steps give the
    frame branch consistent predictions and segment counts of the order of the ground truth (8 per video), which is what real
    checkpoints look like.  Returns frames/s of the resident forward and the segment counts; restores nothing (last thing run)."""
    import copy
    from fact_clip_b200.loss import MatchCriterion
    from fact_clip_b200.utils.synth import make_video
    tcfg = copy.deepcopy(cfg)
    tcfg.Loss.merge(dict(match='o2o', nullw=0.1, bgw=1.0, pc=0.2, a2fc=1.0, sw=5.0))
    net.mcriterion = MatchCriterion(tcfg, N_CLASSES, [])
    net.train()
    net.train_graphs = True
    opt = torch.optim.Adam(net.parameters(), lr=3e-4)
    nb, Tt = 4, 2048                                     # 4 videos of 2048 frames per step: seconds, not minutes
    pool = [[make_video(Tt, IN_DIM, N_CLASSES, seed=7000 + 4 * k + i) for i in range(nb)] for k in range(8)]
    pool = [([v[0].to(dev) for v in b_], [v[1].to(dev) for v in b_]) for b_ in pool]
    t0 = time.perf_counter()
    first = last = None
    for it in range(args.calibrate_steps):
        xs, ys = pool[it % len(pool)]
        opt.zero_grad(set_to_none=True)
        loss, _ = net(xs, ys, compute_loss=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
        opt.step()
        if it == 0:
            first = float(loss.detach())
    last = float(loss.detach())
    torch.cuda.synchronize()
    train_s = time.perf_counter() - t0
    net.eval()
    eng = net.engine()
    for _ in range(3):
        eng.run_packed_graphed(x, ln, lengths)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stash = eng.run_packed_graphed(x, ln, lengths)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    nseg = [st['nseg'].tolist() for st in stash['blocks'] if 'nseg' in st]
    # the same loop with consecutive steps on args.lanes streams / arenas / graphs (how the headline `value` is measured)
    ms_lanes = None
    if args.lanes > 1:
        streams = [torch.cuda.Stream() for _ in range(args.lanes)]

        def lane_step(i):
            k = i % len(streams)
            eng.lane = k
            streams[k].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(streams[k]):
                eng.run_packed_graphed(x, ln, lengths)
            eng.lane = 0

        for i in range(2 * len(streams)):
            lane_step(i)
        for s_ in streams:
            torch.cuda.current_stream().wait_stream(s_)
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            lane_step(i)
        for s_ in streams:
            torch.cuda.current_stream().wait_stream(s_)
        e1.record()
        torch.cuda.synchronize()
        ms_lanes = e0.elapsed_time(e1) / args.steps
    return {'label': 'CALIBRATED weights (not the headline): same forward workload after a short training run with this repo\'s own training step',
            'training': {'steps': args.calibrate_steps, 'videos_per_step': nb, 'frames_per_video': Tt, 'loss_first': first, 'loss_last': last,
                         'seconds': train_s, 'optimizer': 'Adam lr 3e-4, clip_grad_norm 10'},
            'value': B * T / ((ms_lanes or ms) * 1e-3), 'unit': UNIT, 'ms_per_step': ms_lanes or ms,
            'launch': f'one CUDA graph per step, consecutive steps on {args.lanes} streams (like the headline value)' if ms_lanes else 'one CUDA graph per step on one stream',
            'single_stream': {'value': B * T / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms},
            'segments_per_U_block': [[min(s_), max(s_)] for s_ in nseg]}


def run_train(args):
    """BASELINE config 5: FACT_CLIP training step (train-mode forward + loss + hand-written backward + data-parallel
    gradient all-reduce over NVLink + clip_grad_norm_ + Adam step, scripts/train.py:262-268) on synthetic
    Epic-Kitchens-shape videos: epic-kitchens.yaml hyper-parameters with block iUUU, T = 16384, C = 98, one
    video per GPU per step.  Prints one JSON line (rank 0)."""
    import torch.distributed as dist
    from fact_clip_b200 import _lib, config as C, ops
    from fact_clip_b200.loss import MatchCriterion
    from fact_clip_b200.models.blocks import FACT_CLIP
    from fact_clip_b200.parallel import GradAllReducer
    from fact_clip_b200.utils.synth import make_text_embeddings, make_video

    world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib.check(_lib.load().factk_device_check(), 'device_check')
    T, ncls, D = args.train_frames, 98, IN_DIM
    cfg = C.PRESETS['epic_shape']()
    torch.manual_seed(0)
    net = FACT_CLIP(cfg, D, ncls, make_text_embeddings(ncls))
    net.compute_mode = args.mode
    net = net.to(dev).train()
    net.mcriterion = MatchCriterion(cfg, ncls, [0])
    net.train_graphs = not args.no_graph
    # torch's own Adam (scripts/train.py builds torch.optim.Adam, epic-kitchens.yaml:74-80); fused=True is its single multi-tensor
    # kernel instead of the foreach chain -- same update rule (clip + step 2.0 -> 1.4 ms); FACTK_FUSED_ADAM=0 selects torch's default
    adam_kw = dict(fused=True) if os.environ.get('FACTK_FUSED_ADAM', '1') != '0' else {}
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=0.0, **adam_kw)
    red = GradAllReducer()
    net.grad_ready_hook = red.on_bucket
    nparam = sum(p.numel() for p in net.parameters())
    # one pinned host video per rank and step slot (two alternate), copied to the device inside the step
    # Weak scaling measures the data-parallel machinery, so every rank trains on the SAME two synthetic videos (alternating by
    # step; the channel-dropout masks differ per rank): the data-dependent segment counts -- hence the GRU chain lengths, which
    # dominate the step under random-init weights -- are then identical across ranks and across GPU counts.
    hosts = []
    for k in range(2):
        x, y = make_video(T, D, ncls, seed=9000 + k, nseg=args.train_nseg)
        hosts.append((x.pin_memory(), y.pin_memory()))
    net.train_engine().seed += 7919 * rank
    ev = lambda: torch.cuda.Event(enable_timing=True)
    phases = {'fwd_loss': 0.0, 'bwd': 0.0, 'allreduce_wait': 0.0, 'clip_opt': 0.0}
    info = {}

    # input pipeline: the pinned-host -> device copy of the NEXT step's video is issued on a copy stream while this step computes
    # (a prefetching data loader); every timed step still pays for one 134 MB copy inside the timed region, off the critical path
    copy_stream, staged = torch.cuda.Stream(), {}

    def prefetch(i):
        x, y = hosts[i % 2]
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            xd, yd = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        staged[i % 2] = (xd, yd, done)

    def step(i, timed):
        x, y = hosts[i % 2]
        e = [ev() for _ in range(5)]
        e[0].record()
        if args.no_prefetch:
            xs, ys = [x.to(dev, non_blocking=True)], [y.to(dev, non_blocking=True)]
        else:
            if i % 2 not in staged:
                prefetch(i)
            xd, yd, done = staged.pop(i % 2)
            torch.cuda.current_stream().wait_event(done)
            xd.record_stream(torch.cuda.current_stream())
            yd.record_stream(torch.cuda.current_stream())
            xs, ys = [xd], [yd]
            prefetch(i + 1)
        opt.zero_grad(set_to_none=True)
        loss, saves = net(xs, ys, compute_loss=True)
        e[1].record()
        loss.backward()
        e[2].record()
        info['allreduce_bytes'] = red.finish()
        e[3].record()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)                   # epic-kitchens.yaml: clip_grad_norm 10.0
        opt.step()
        e[4].record()
        info['loss'] = saves[0]['loss']['loss']
        info['nseg'] = [int(st['nseg'][0]) for st in net._last['blocks'] if 'nseg' in st]
        return e

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    ops.COUNTERS['launches'] = 0
    step(0, False)                                   # eager: the kernels launched per step are counted here
    info['launches'] = ops.COUNTERS['launches']
    for i in range(1, W + 2):                        # graph capture happens on the second and third call of a shape
        step(i, False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ops.COUNTERS['launches'] = 0
    e0, e1 = ev(), ev()
    e0.record()
    evs = [step(W + i, True) for i in range(args.steps)]
    e1.record()
    barrier()
    clocks = sampler.summary()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if args.profile and rank == 0:
        net.train_graphs = False
        step(W + args.steps, False)          # eager warm-up of the un-graphed path
        ops.TIMER = ops.KernelTimer(None)
        t0 = time.perf_counter()
        step(W + args.steps, False)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        prof = ops.TIMER.collect(skip_steps=0, steps=1)
        ops.TIMER = None
        tot = sum(v['ms'] for v in prof.values())
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
            sys.stderr.write(f"  {k:28s} n={v['n']:5d} {v['ms']:9.3f} ms {100 * v['ms'] / tot:5.1f}%\n")
        sys.stderr.write(f"  sum of kernel times {tot:.2f} ms; wall of the profiled step {wall:.1f} ms\n")
    if args.timeline and rank == 0:                  # device timeline of one graph-replayed step: busy time, idle gaps
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step(W + args.steps, False)
            torch.cuda.synchronize()
        ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                    key=lambda t: t[0])
        busy, gaps, end = 0.0, [], ks[0][0]
        for a, b, name in ks:
            if a > end:
                gaps.append((a - end, name))
            busy += max(0.0, b - max(a, end))
            end = max(end, b)
        span = end - ks[0][0]
        sys.stderr.write(f"  timeline: {len(ks)} device activities, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms\n")
        for gdur, name in sorted(gaps, reverse=True)[:12]:
            sys.stderr.write(f"    gap {gdur / 1e3:7.3f} ms before {name[:90]}\n")
        small = sum(gd for gd, _ in gaps if gd < 20)
        sys.stderr.write(f"    gaps < 20 us: {sum(1 for gd, _ in gaps if gd < 20)} totalling {small / 1e3:.2f} ms\n")
        agg = {}
        for a, b, name in ks:
            k = name.split('(')[0].split('<')[0][-60:]
            if 'at::native' in name:
                import re as _re
                m_ = _re.search(r'(\w+Functor|\w+Ops?\b|\w+_kernel_cuda|BinaryFunctor<[^>]*>|\w+Op<[^>]*>)', name.split('at::native::', 2)[-1][20:])
                k = ('aten:' + name.split('at::native::')[1][:26] + ':' + (m_.group(1) if m_ else ''))[:60]
            agg[k] = (agg.get(k, (0, 0))[0] + (b - a), agg.get(k, (0, 0))[1] + 1)
        for k, (d, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
            sys.stderr.write(f"    {k:60s} n={n:5d} {d / 1e3:8.3f} ms\n")
    for e in evs:
        for k, (a, b) in zip(phases, zip(e[:-1], e[1:])):
            phases[k] += a.elapsed_time(b) / args.steps
    frames = T * world * args.steps
    res = {
        'metric': 'frames/sec FACT_CLIP training step (T=16384, 2048-d, data parallel)', 'value': frames / (ms * 1e-3), 'unit': UNIT,
        'n_gpus': world, 'steps': args.steps, 'warmup': W, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': args.mode if args.mode == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': f'FACT_CLIP training step, epic-kitchens.yaml shape (block iUUU, MSTCN++, F=A=256, M=300, fpos, o2m matching), '
                               f'T={T}, D={D}, C={ncls}, one video per GPU per step; train-mode forward + loss + backward + DP all-reduce + '
                               f'clip_grad_norm_ + Adam step; random-init weights (segments per U block below)',
                   'segments_per_U_block_rank0': info.get('nseg'), 'params': nparam, 'loss_rank0': info.get('loss'),
                   'l2_policy': f'inputs larger than L2 ({T * D * 4 / 2**20:.0f} MiB of fp32 features per step; activations far larger)',
                   'input_pipeline': 'per-step copy at the start of the step' if args.no_prefetch else
                                     'next step\'s video copied pinned-host -> device on a copy stream during this step (one copy per timed step)'},
        'clocks': clocks, 'gpu_launches': info.get('launches'), 'launch': 'cuda-graph' if net.train_graphs else 'eager',
        'e2e': {'value': frames / (ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': (T * D * 4 + T * 8) * world, 'd2h_bytes_per_step': T * 8 * world,
                'note': 'the timed step includes the pinned-host -> device copy of the video and the device -> host copy of the predictions and the loss'},
        'phases_ms_rank0': phases,
        'allreduce': {'bytes_per_step_per_rank': info.get('allreduce_bytes'), 'buckets': len(cfg.FACT.block) + 1,
                      'overlap': 'each section (CLIP head, then blocks last to first) is all-reduced in place as soon as the backward pass leaves it',
                      'wait_ms_after_backward': phases['allreduce_wait']},
    }
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--videos', type=int, default=64, help='videos per GPU per step')
    ap.add_argument('--mode', default='bf16')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-split', type=int, default=4, help='sub-batches per step on the end-to-end (pipelined) path')
    ap.add_argument('--no-graph', action='store_true', help='eager kernel launches instead of one CUDA graph per batch')
    ap.add_argument('--resident-only', action='store_true', help='stop after the resident loop (profiling aid)')
    ap.add_argument('--lanes', type=int, default=3, help='resident measurement: steps alternate between this many streams / activation arenas')
    ap.add_argument('--no-prefetch', action='store_true', help='(--train) copy each step\'s video to the device at the start of the step instead of during the previous step')
    ap.add_argument('--timeline', action='store_true', help='(--train) torch.profiler device timeline of one step: busy / idle / top kernels, to stderr')
    ap.add_argument('--profile', action='store_true', help='print a CUDA-event breakdown per kernel family to stderr')
    ap.add_argument('--train', action='store_true', help='BASELINE config 5: data-parallel training step (Epic-Kitchens shape, T=16384)')
    ap.add_argument('--calibrate-steps', type=int, default=600, help='training steps for the second (calibrated-weights) regime; 0 skips it')
    ap.add_argument('--train-frames', type=int, default=16384)
    ap.add_argument('--train-nseg', type=int, default=52, help='ground-truth segments per synthetic training video (epic o2m average)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.train:
        return run_train(args)

    import torch.distributed as dist
    from fact_clip_b200 import _lib, ops
    from fact_clip_b200.utils.synth import make_video

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    from fact_clip_b200.parallel import bind_to_gpu_numa_node
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)      # before the pinned feature buffers exist
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib.check(_lib.load().factk_device_check(), 'device_check')

    net, cfg = build_model(args.mode)
    net = net.to(dev)
    eng = net.engine()
    eng.use_graph = not args.no_graph
    B, T = args.videos, T_FRAMES
    # rank r owns videos {r*B .. r*B+B-1} of the sweep (videos shard, no data-path collective)
    host = torch.empty(B, T, IN_DIM, pin_memory=True)
    for b in range(B):
        host[b].copy_(make_video(T, IN_DIM, N_CLASSES, seed=5000 + rank * B + b)[0])
    x = host.to(dev)
    lengths = [T] * B
    ln = torch.tensor(lengths, dtype=torch.int32, device=dev)
    seq_host = [host[b] for b in range(B)]
    labels = [torch.zeros(T, dtype=torch.long) for _ in range(B)]

    lanes = [torch.cuda.Stream() for _ in range(args.lanes)] if args.lanes > 1 else None
    lane_no = [0]

    def step_resident():
        if lanes is None or ops.TIMER is not None:
            return eng.run_packed_graphed(x, ln, lengths)
        # `--lanes N`: consecutive steps (independent batches) alternate between N streams, each with its own activation arena and
        # CUDA graph, so one step's latency-bound stretches (the GRU chain) overlap the next step's tensor-core work
        k = lane_no[0] = (lane_no[0] + 1) % len(lanes)
        eng.lane = k
        lanes[k].wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(lanes[k]):
            out = eng.run_packed_graphed(x, ln, lengths)
        eng.lane = 0
        return out

    def drain_lanes():
        if lanes is not None:
            for s_ in lanes:
                torch.cuda.current_stream().wait_stream(s_)

    pending = []
    e2e_split = [args.e2e_split]

    def step_e2e():
        # public pipelined API: this step's H2D copy (pinned host -> device, side stream) overlaps the previous
        # step's kernels; every step's copy-in and prediction copy-out happen inside the timed region
        # The step's videos go through in `e2e_split` sub-batches: the path is PCIe-bound, and smaller sub-batches shorten the
        # un-overlapped tail (the last sub-batch's kernels) without changing what is copied or computed per step.
        n = max(B // e2e_split[0], 1)
        for k in range(0, B, n):
            pending.append(net.submit(seq_host[k:k + n], labels[k:k + n]))
            if len(pending) > 1:
                pending.pop(0).result()

    def drain_e2e():
        while pending:
            pending.pop(0).result()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, drain=None):
        for _ in range(warmup):
            fn()
        if drain:
            drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain:
            drain()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    sampler = ClockSampler(local)
    ops.COUNTERS['launches'] = 0
    sampler.start()
    ms = timed(step_resident, args.steps, max(args.warmup, 3), drain=drain_lanes)
    ms_single = timed(lambda: eng.run_packed_graphed(x, ln, lengths), args.steps, max(args.warmup, 3)) if lanes is not None else ms
    def profile_pass():
        ops.TIMER = ops.KernelTimer(None)
        for _ in range(3):
            step_resident()
        prof = ops.TIMER.collect(skip_steps=1, steps=2)
        ops.TIMER = None
        tot = sum(v['ms'] for v in prof.values())
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
            sys.stderr.write(f"  {k:28s} n/step={v['n'] // 2:4d} {v['ms'] / 2:8.3f} ms/step {100 * v['ms'] / tot:5.1f}%\n")
        sys.stderr.write(f"  total {tot / 2:.3f} ms/step (sum of kernel times)\n")

    if args.resident_only:            # profiling aid (ncu launch lists): nothing but the resident loop
        if args.profile:
            profile_pass()
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'value': B * T * world * args.steps / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms / args.steps,
                              'gpu_launches': eng.last_launches if eng.use_graph else None, 'note': '--resident-only'}))
        return
    clocks = sampler.summary()
    launches = eng.last_launches if eng.use_graph else ops.COUNTERS['launches'] // (args.steps + max(args.warmup, 3))
    # kernel-level pass: the same steps with eager launches and CUDA events around every launch of the dominant kernel
    ops.TIMER = ops.KernelTimer(('tcn_layer', 'tcn_conv3', 'tcn_1x1', 'clip', 'in_proj', 'conv_out'))
    ms_eager = timed(step_resident, args.steps, 1)
    ktimes = ops.TIMER.collect(skip_steps=1, steps=args.steps)
    ops.TIMER = None
    if args.profile:
        profile_pass()
    # the ceiling of the host-fed number: all ranks copy their pinned feature buffer to the device at the same time, nothing else
    xin = torch.empty_like(x)
    def step_h2d():
        xin.copy_(host, non_blocking=True)
    ms_h2d = timed(step_h2d, 4, 2)
    h2d_ceiling = B * T * IN_DIM * 4 * world * 4 / (ms_h2d * 1e-3) / 1e9          # GB/s aggregate over the ranks
    del xin
    ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3), drain=drain_e2e)
    # same pipeline with the features stored as bf16 on the host (input staging): half the PCIe bytes
    host16 = host.to(torch.bfloat16).pin_memory()
    seq_host = [host16[b] for b in range(B)]
    e2e_split[0] = max(args.e2e_split // 2, 1)     # half the copy time per video: fewer, larger sub-batches keep it copy-bound
    ms_e2e16 = timed(step_e2e, args.steps, max(args.warmup, 3), drain=drain_e2e)

    frames_step = B * T * world
    value = frames_step * args.steps / (ms * 1e-3)
    e2e = frames_step * args.steps / (ms_e2e * 1e-3)

    # roofline of the dominant kernel family: the dilated residual layer of the frame branch
    pk = peaks()
    F = cfg.Bi.f_dim
    n_layers = sum(b_.f_layers for b_ in (cfg.Bi, cfg.Bu, cfg.BU, cfg.BU))
    tcn_ms = sum(ktimes[k]['ms'] for k in ('tcn_layer', 'tcn_conv3', 'tcn_1x1'))
    t_layer_ms = tcn_ms / max(ktimes['tcn_layer']['n'] + ktimes['tcn_conv3']['n'], 1)
    ach_tf = alg_flops_tcn_layer(F) * B * T / (t_layer_ms * 1e-3) / 1e12 if t_layer_ms > 0 else 0.0
    stash = step_resident()
    nseg = [st['nseg'].tolist() for st in stash['blocks'] if 'nseg' in st]
    # FACT_CLIP logit head (blocks.py:161-175, 822-826): its three GEMMs (437(+75 zero columns)->512, 512->512, 512->C) against the tensor peak
    Ppj = cfg.CLIP.projection_hidden_dim
    clip_flops = 2.0 * (512 * Ppj + Ppj * 512 + 512 * N_CLASSES) * B * T
    clip_ms = ktimes['clip']['ms'] / max(args.steps, 1)
    clip_tf = clip_flops / (clip_ms * 1e-3) / 1e12 if clip_ms > 0 else 0.0
    hbm_of = lambda tag, bytes_per_frame: (bytes_per_frame * B * T / (ktimes[tag]['ms'] / max(ktimes[tag]['n'], 1) * 1e-3) / 1e9 / pk['hbm']
                                            if ktimes[tag]['n'] else None)

    res = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': args.mode if args.mode == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': f'FACT_CLIP {PRESET} forward, T={T}, D={IN_DIM}, C={N_CLASSES}, F=A=256, M=75, block iuUU; '
                               f'{B} videos/GPU/step, random-init weights, segment-structured synthetic features',
                   'videos_per_gpu': B, 'frames_per_step': frames_step,
                   'launch': 'cuda-graph' if eng.use_graph else 'eager',
                   'resident_lanes': args.lanes,
                   'resident_lanes_note': 'consecutive steps (independent batches) alternate between this many streams, each with its own '
                                          'activation arena and CUDA graph: one step\'s latency-bound GRU chain overlaps the next step\'s '
                                          'tensor-core kernels; single_stream below is the same loop on one stream',
                   'l2_policy': f'inputs larger than L2 ({B * T * IN_DIM * 4 / 2**20:.0f} MiB of fp32 features per step)',
                   'segments_per_U_block_rank0': [[min(s), max(s)] for s in nseg], 'numa_rank0': numa},
        'clocks': clocks, 'gpu_launches': launches,
        'single_stream': {'value': frames_step * args.steps / (ms_single * 1e-3), 'unit': UNIT, 'ms_per_step': ms_single / args.steps},
        'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': B * T * IN_DIM * 4 * world,
                'd2h_bytes_per_step': B * T * 8 * world, 'ms_per_step': ms_e2e / args.steps,
                'features': 'fp32 on the host (the reference format)', 'sub_batches_per_step': args.e2e_split,
                'h2d_gbs': B * T * IN_DIM * 4 * world * args.steps / (ms_e2e * 1e-3) / 1e9,
                'h2d_ceiling_gbs': h2d_ceiling,
                'frac_of_h2d_ceiling': (B * T * IN_DIM * 4 * world * args.steps / (ms_e2e * 1e-3) / 1e9) / h2d_ceiling,
                'h2d_ceiling_note': 'aggregate pinned host -> device copy rate of the same buffers with all ranks copying at once and no '
                                    'kernels running, measured in this run; the end-to-end path is bound by it (PCIe / host memory), not by a kernel'},
        'e2e_bf16_features': {'value': frames_step * args.steps / (ms_e2e16 * 1e-3), 'unit': UNIT,
                              'h2d_bytes_per_step': B * T * IN_DIM * 2 * world, 'd2h_bytes_per_step': B * T * 8 * world,
                              'ms_per_step': ms_e2e16 / args.steps,
                              'features': 'pre-converted to bf16 on the host (input staging option; not the headline)',
                              'sub_batches_per_step': max(args.e2e_split // 2, 1)},
        'roofline': {'bound': 'tensor', 'kernel': 'tcn_layer_kernel: fused dilated residual layer (conv3+ReLU+1x1+residual), 40 launches per forward',
                     'achieved': ach_tf, 'peak': pk['tf_sust'], 'unit': 'TFLOP/s', 'frac': ach_tf / pk['tf_sust'],
                     'traffic': TCN_DRAM_BYTES_PER_LAUNCH.get(B), 'traffic_source': 'profile constant, not measured live: profiles/r1_tcn_layer_ncu_full.txt (dram read + write per launch, ncu --set full)',
                     'peak_source': pk['src'] + ' (sustained bf16, kernel timed inside a long step)',
                     'ms_per_layer': t_layer_ms, 'share_of_step': tcn_ms / args.steps / (ms_eager / args.steps),
                     # both one-sided fractions of the same launch (algorithmic bytes: x in + y out, bf16)
                     'one_sided': {'tensor': ach_tf / pk['tf_sust'],
                                   'hbm': (4.0 * F * B * T / (t_layer_ms * 1e-3) / 1e9 / pk['hbm']) if t_layer_ms > 0 else 0.0,
                                   'hbm_peak_gbs': pk['hbm']},
                     'timed_in': 'eager pass of the same steps (CUDA events around each launch); the headline loop replays one CUDA graph per step',
                     # 34.6 MFLOP of algorithmic work per frame at this config (segment-count dependent terms
                     # excluded) -> the whole forward against the same tensor peak, per GPU
                     'whole_forward': {'flops_per_frame': 34.6e6, 'achieved': value / world * 34.6e6 / 1e12, 'unit': 'TFLOP/s',
                                       'frac': value / world * 34.6e6 / 1e12 / pk['tf_sust']}},
    }
    res['clip_head'] = {'kernels': 'three tcgen05 GEMMs (CTA-pair bf16 x2, bf16 text GEMM with 1/temp) + LayerNorm/ReLU and L2-normalise row kernels',
                        'gemm_ms_per_step': clip_ms, 'flops_per_step': clip_flops, 'achieved': clip_tf, 'unit': 'TFLOP/s',
                        'one_sided': {'tensor': clip_tf / pk['tf_sust']}, 'peak': pk['tf_sust'],
                        'note': 'GEMM launches only (tag clip), timed with CUDA events in the eager pass; padded first layer counted at K = 512'}
    res['hbm_kernels'] = {'in_proj': {'bytes_per_frame': 4 * IN_DIM + 2 * F, 'frac_of_hbm': hbm_of('in_proj', 4 * IN_DIM + 2 * F)},
                          'conv_out': {'bytes_per_frame': 2 * F + 2 * 512, 'frac_of_hbm': hbm_of('conv_out', 2 * F + 2 * 512)},
                          'hbm_peak_gbs': pk['hbm'], 'ncu': 'profiles/r2_frame_gemms_ncu.md (dram bytes per launch)'}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)       # the CPU baseline may use every host core
            threads = os.cpu_count() or 1
            n_cpu = 24      # ~10-30 s of CPU work on the box's 16 cores
            fps, dt, kind = cpu_baseline(n_cpu, threads)
            try:
                par = parity_vs_reference(net) if kind == 'reference' else None
            except Exception as e:
                par = {'error': f'{type(e).__name__}: {e}'}
            if par is not None:
                res['config']['parity_vs_reference'] = par
            res['cpu_baseline'] = {'value': fps, 'unit': UNIT, 'cores': threads, 'kind': kind,
                                   'sample': f'{n_cpu} videos x T={T} of the same workload, '
                                             + ('stock reference FACT_CLIP from baseline/_ref' if kind == 'reference' else 'oracle port')
                                             + f' (torch CPU fp32, eval, no_grad), {dt:.1f} s'}
    # the calibrated regime trains the net in place: it runs LAST (the parity check above compares against the reference's random init)
    if world == 1 and args.calibrate_steps > 0:
        try:
            res['calibrated_regime'] = calibrated_regime(net, cfg, dev, x, ln, lengths, B, T, args)
        except Exception as e:          # the second regime never takes the headline line down
            import traceback
            res['calibrated_regime'] = {'error': f'{type(e).__name__}: {str(e)[:200]}', 'where': traceback.format_exc()[-1500:]}
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
