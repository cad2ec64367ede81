"""Loss value + matching on device at the bench workload (HAViD CLIP config, 64 videos x 4096 frames): time of the eager
forward alone, of forward + loss, the split of the loss part into device kernels and host matching, and the CPU oracle
(oracle/loss_oracle.py on forward tensors copied to the host) on a sample of the same videos.  Wall-clock with
synchronisation on both sides: the loss path has two small host round-trips by design."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from fact_clip_b200 import config as C, ops  # noqa: E402
from fact_clip_b200.models.blocks import FACT_CLIP  # noqa: E402
from fact_clip_b200.models.loss import MatchCriterion  # noqa: E402
from fact_clip_b200.utils.synth import make_batch, make_text_embeddings  # noqa: E402


def wall(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    epic = os.environ.get('PRESET', 'havid') == 'epic'      # BASELINE config 5 shape: T = 16384, 300 tokens, o2m matching
    videos, T, ncls = int(os.environ.get('VIDEOS', 2 if epic else 64)), (16384 if epic else 4096), (98 if epic else 75)
    cfg = C.PRESETS['epic_shape' if epic else 'havid_view0_lh_pt_holdout']()
    cfg.merge(dict(Loss=dict(nullw=0.05, bgw=0.5)))
    torch.manual_seed(0)
    net = FACT_CLIP(cfg, 2048, ncls, make_text_embeddings(ncls)).eval()
    net.compute_mode = 'bf16'
    net = net.cuda()
    net.mcriterion = MatchCriterion(cfg, ncls, [0])
    xs, ys = make_batch([T] * videos, 2048, ncls, base_seed=7, nseg=52 if epic else 8)
    xd, yd = [x.cuda() for x in xs], [y.cuda() for y in ys]
    for _ in range(2):
        net(xd, yd, compute_loss=True)
    fwd = wall(lambda: net(xd, yd), 5)
    both = wall(lambda: net(xd, yd, compute_loss=True), 5)
    n0 = ops.COUNTERS['launches']
    loss, saves = net(xd, yd, compute_loss=True)
    launches = ops.COUNTERS['launches'] - n0
    # host share of the loss: matching of the [M,S] costs
    from fact_clip_b200 import loss as Lm
    t0 = time.perf_counter()
    rng = np.random.default_rng(0)
    for _ in range(videos):
        Lm.assign(rng.standard_normal((300, 52) if epic else (75, 8)).astype(np.float32), np.arange(52 if epic else 8) % 40, cfg.Loss.match)
    host_match = (time.perf_counter() - t0) * 1e3
    line = dict(workload=f'{"epic shape" if epic else "havid holdout"} FACT_CLIP, {videos} videos x {T} frames, eager launches',
                forward_ms=round(fwd, 2), forward_plus_loss_ms=round(both, 2), loss_ms=round(both - fwd, 2),
                host_matching_ms=round(host_match, 2), launches_forward_plus_loss=launches, loss=float(loss),
                finite=bool(np.isfinite(float(loss))))
    # CPU baseline: the oracle's loss on the same tensors (forward NOT included), a sample of videos
    if os.environ.get('CPU', '1') == '1':
        import fact_oracle as O
        import loss_oracle as LO
        hp = O.hparams_from_cfg(cfg, 2048, ncls)
        lp = LO.loss_params(cfg, bg_ids=[0])
        text = net.text_embeddings.cpu()
        net.keep_attn = True
        net(xd, yd)
        n = min(4, videos)
        t_cpu = 0.0
        diffs = []
        for b in range(n):
            net.stash_video(b)
            out = dict(blocks=[], projected_frame_embeddings=net.projected_frame_embeddings[:, 0].float().cpu())
            for blk in net.block_list:
                st = dict(frame_clogit=blk.frame_clogit[:, 0].float().cpu(), action_clogit=blk.action_clogit[:, 0].float().cpu())
                if hasattr(blk, 'a2f_attn_logit'):
                    st.update(f2a_attn_logit=blk.f2a_attn_logit[0].float().cpu(), a2f_attn_logit=blk.a2f_attn_logit[0].float().cpu(),
                              a2f_attn=blk.a2f_attn[0].float().cpu())
                if hasattr(blk, 'seg_clogit') and hasattr(blk, 'tdu'):
                    st.update(seg_clogit=blk.seg_clogit[:, 0].float().cpu(), seg_label=blk.tdu.seg_label.cpu(),
                              seg_lens=blk.tdu.seg_lens.cpu())
                out['blocks'].append(st)
            t0 = time.perf_counter()
            ref = LO.loss_video(out, hp, ys[b], lp, text_embeddings=text)
            t_cpu += time.perf_counter() - t0
            diffs.append(abs(float(ref['loss']) - saves[b]['loss']['loss']) / max(1.0, abs(float(ref['loss']))))
        line.update(cpu_oracle_loss_ms_per_video=round(t_cpu / n * 1e3, 1), cpu_sample_videos=n,
                    cpu_oracle_loss_ms_for_batch=round(t_cpu / n * 1e3 * videos, 1),
                    max_rel_diff_vs_oracle_on_same_tensors=max(diffs))
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
