"""Concurrent pinned host -> device copy bandwidth over sets of GPUs of one node: the ceiling of the end-to-end (host-fed)
frames/s at N GPUs (8 KiB of fp32 features per frame).

    python tools/h2d_bw_multi.py --sets 0 0,1 0,1,2,3 0,2,4,6 0,1,2,3,4,5,6,7

One process, one thread + stream per GPU; every GPU copies its own 1 GiB pinned buffer `reps` times, all starting together;
reports per-set aggregate and per-GPU GiB/s as one JSON line per set."""
import argparse
import json
import threading
import time

import torch


def run_set(gpus, mib=1024, reps=6):
    bufs = []
    for g in gpus:
        with torch.cuda.device(g):
            h = torch.empty(mib * 2**20 // 4, dtype=torch.float32, pin_memory=True)
            d = torch.empty_like(h, device=f'cuda:{g}')
            d.copy_(h, non_blocking=True)
            bufs.append((g, h, d, torch.cuda.Stream(device=g)))
    for g, *_ in bufs:
        torch.cuda.synchronize(g)
    barrier = threading.Barrier(len(gpus) + 1)
    times = {}

    def work(g, h, d, st):
        with torch.cuda.device(g), torch.cuda.stream(st):
            barrier.wait()
            t0 = time.perf_counter()
            for _ in range(reps):
                d.copy_(h, non_blocking=True)
            st.synchronize()
            times[g] = time.perf_counter() - t0
    th = [threading.Thread(target=work, args=b) for b in bufs]
    for t in th:
        t.start()
    barrier.wait()
    for t in th:
        t.join()
    per = {g: reps * mib / 1024 / times[g] for g in gpus}
    return dict(gpus=gpus, aggregate_gib_s=round(reps * mib / 1024 * len(gpus) / max(times.values()), 1),
                per_gpu_gib_s={str(g): round(v, 1) for g, v in per.items()})


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--sets', nargs='+', default=['0'])
    a = ap.parse_args()
    n = torch.cuda.device_count()
    for s in a.sets:
        gpus = [int(x) for x in s.split(',')]
        if max(gpus) >= n:
            continue
        print(json.dumps(run_set(gpus)))
