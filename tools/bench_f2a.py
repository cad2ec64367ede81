"""Time the f2a direction of X2Y_map at the metric shape: fused tcgen05 kernel (f2a_fused.cu) against the unfused chain it
replaces (tcgen05 logit GEMM -> column statistics -> mma.sync apply -> combine).  CUDA events, L2 flushed between iterations.

    python tools/bench_f2a.py [B slot M H]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fact_clip_b200 import ops  # noqa: E402
from fact_clip_b200.ops import S  # noqa: E402


def main():
    B, slot, M, H = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 4096, 75, 512)
    dev, bf = 'cuda', torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    rows = torch.randn(B, slot, H, device=dev, generator=g).to(bf)
    qt = (torch.randn(B, M, H, device=dev, generator=g) * 0.1).to(bf)
    cb = torch.randn(B, M, device=dev, generator=g)
    ln = torch.full((B,), slot, dtype=torch.int32, device=dev)
    Mp = (M + 3) // 4 * 4
    logit = torch.empty(B, slot, Mp, device=dev)
    out_a, out_b = torch.empty(B, M, H, device=dev), torch.empty(B, M, H, device=dev)
    ws_a = torch.empty(ops.col_softmax_ws(B, slot, M, H), device=dev)
    ws_b = torch.empty(ops.f2a_fused_ws(B, slot, M, H), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def chain():
        ops.gemm([S(rows, qt)], M, logit, len=ln, bias=cb, tc=True)
        ops.col_softmax_apply(logit, rows, out_a, M, ws_a, len=ln, E=H)

    def fused():
        ops.f2a_fused(rows, qt, out_b, M, ws_b, len=ln)

    res = {}
    for name, fn in (('chain', chain), ('fused', fused)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        res[name] = ts[len(ts) // 2]
    # stage timeline of CTA (0, 0): softmax warp 4, lane 0 (clock64)
    from fact_clip_b200 import _lib
    dbg = torch.zeros(64, dtype=torch.int64, device=dev)
    _lib.load().factk_f2a_debug(dbg.data_ptr())
    fused()
    torch.cuda.synchronize()
    _lib.load().factk_f2a_debug(None)
    d = dbg.tolist()
    t0 = d[0]
    print(f'timeline (cycles from CTA entry): setup done {d[1] - t0}, tiles done {d[2] - t0}, o_full {d[3] - t0}, exit {d[4] - t0}')
    for t in range(12):
        a = d[8 + 4 * t: 12 + 4 * t]
        if a[0]:
            print(f'  tile {t}: wait S {a[1] - a[0]:6d}  softmax {a[2] - a[1]:6d}  wait P free {a[3] - a[2]:6d}   (start {a[0] - t0})')
    err = float((out_a - out_b).norm() / out_a.norm())
    bytes_rows = B * slot * H * 2
    print(f'B={B} slot={slot} M={M}: chain {res["chain"]:.1f} us, fused {res["fused"]:.1f} us '
          f'({bytes_rows / res["fused"] / 1e3:.0f} GB/s of rows), fused vs chain rel-L2 {err:.2e}')


if __name__ == '__main__':
    main()
