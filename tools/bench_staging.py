"""From .npy files to predictions: the dataset -> device staging of fact_clip_b200/staging.py on the bench workload
(HAViD CLIP config, 64 videos x 4096 frames x 2048-d fp32 stored TRANSPOSED as (D, T) like the reference's `feature.T`
datasets).  Reports the host-side staging rate alone (files in the page cache -> pinned arenas) and the pipelined sweep
(staging + H2D + forward), next to the reference's way of doing it (np.load -> .T -> astype -> from_numpy -> .cuda())."""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fact_clip_b200 import config as C  # noqa: E402
from fact_clip_b200.models.blocks import FACT_CLIP  # noqa: E402
from fact_clip_b200.staging import FeatureStager, run_sweep  # noqa: E402
from fact_clip_b200.utils.synth import make_text_embeddings, make_video  # noqa: E402


def main():
    videos, T, D, ncls = int(os.environ.get('VIDEOS', 64)), 4096, 2048, 75
    d = tempfile.mkdtemp(dir=os.environ.get('STAGE_DIR', '/dev/shm'))
    try:
        names = [f'v{i:03d}' for i in range(videos)]
        for i, n in enumerate(names):
            x, _ = make_video(T, D, ncls, seed=500 + i, nseg=8)
            np.save(os.path.join(d, n + '.npy'), np.ascontiguousarray(x.numpy().T))          # (D, T) on disk
        cfg = C.PRESETS['havid_view0_lh_pt_holdout']()
        torch.manual_seed(0)
        net = FACT_CLIP(cfg, D, ncls, make_text_embeddings(ncls)).eval()
        net.compute_mode = 'bf16'
        net = net.cuda()
        frames = videos * T
        line = dict(workload=f'{videos} videos x {T} x {D} fp32 .npy files stored (D, T), page cache (tmpfs)', frames=frames)
        for workers in (4, 16):
            st = FeatureStager(d, names, transpose=True, batch_videos=16, workers=workers, depth=3)
            for b in st:
                b.release()                                     # warm: arenas allocated and pinned
            t0 = time.perf_counter()
            for b in st:
                b.release()
            dt = time.perf_counter() - t0
            line[f'staging_only_GBps_workers{workers}'] = round(frames * D * 4 / dt / 1e9, 2)
            st.close()
        st = FeatureStager(d, names, transpose=True, batch_videos=16, workers=16, depth=3, device_transpose=True)
        for b in st:
            b.release()
        t0 = time.perf_counter()
        for b in st:
            b.release()
        line['staging_only_GBps_file_layout_workers16'] = round(frames * D * 4 / (time.perf_counter() - t0) / 1e9, 2)
        st.close()
        for dtype, key, dev_t in ((torch.float32, 'sweep_fp32', False), (torch.bfloat16, 'sweep_bf16_arena', False),
                                  (torch.float32, 'sweep_fp32_device_transpose', True)):
            st = FeatureStager(d, names, transpose=True, batch_videos=16, workers=16, depth=3, dtype=dtype, device_transpose=dev_t)
            for _ in run_sweep(net, st):
                pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 0
            for _, saves in run_sweep(net, st):
                n += len(saves)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            line[key + '_frames_per_s'] = round(frames / dt)
            st.close()
        # the reference's way (utils/dataset.py:12-21 + run_eval.py:32-33), feeding the same forward one batch at a time
        t0 = time.perf_counter()
        for i in range(0, videos, 16):
            seqs = []
            for n in names[i:i + 16]:
                f = np.load(os.path.join(d, n + '.npy')).T
                f = f.astype(np.float32) if f.dtype != np.float32 else f
                seqs.append(torch.from_numpy(f).cuda())
            net(seqs, None)
        torch.cuda.synchronize()
        line['reference_style_loading_frames_per_s'] = round(frames / (time.perf_counter() - t0))
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == '__main__':
    main()
