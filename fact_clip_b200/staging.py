"""Dataset -> device input staging: the step BEFORE the forward.

The reference loads every video with ``np.load`` -> optional ``feature.T`` (a non-contiguous view) -> ``astype(float32)`` ->
``torch.from_numpy`` -> a pageable ``.cuda()`` per video on the training thread (utils/dataset.py:12-21, 82-131;
scripts/run_eval.py:32-33).  At the forward rates of this package (millions of frames/s, i.e. tens of GB/s of fp32 features)
that path is the bottleneck, so this module does the same work off the critical path:

  * ``FeatureStager`` memory-maps the ``.npy`` files and writes each video -- transposed and cast in ONE pass -- straight into
    a pinned host arena (a ring of ``depth`` arenas, filled by a small thread pool; numpy's copy loops release the GIL);
  * the features can be stored as bf16 in the arena (half the PCIe bytes; the engine takes bf16 rows in bf16 compute mode);
  * ``run_sweep`` feeds the batches to ``net.submit()`` (double-buffered async H2D + CUDA-graph replay) and recycles an arena
    only after the batch that used it has finished.

Same values as the reference's ``load_feature`` (checked in tests/test_staging.py); no reference code is used.
"""
import os
import queue
import threading
import warnings
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

# read-only memory maps make torch.from_numpy warn; the stager only reads them.  Installed ONCE at import: the filter list
# is process-global and warnings.catch_warnings() is not thread-safe inside the pool threads.
warnings.filterwarnings('ignore', message='The given NumPy array is not writ', category=UserWarning)


def load_feature(feature_dir, video, transpose):
    """The reference's ``load_feature`` semantics (utils/dataset.py:12-21) as one array: (T, D) float32."""
    a = np.load(os.path.join(feature_dir, video + '.npy'), mmap_mode='r')
    if transpose:
        a = a.T
    return np.ascontiguousarray(a, dtype=np.float32)


class Batch:
    """One staged batch: ``names``, ``seqs`` (views of a pinned arena, (T_i, D) each) and the arena slot to give back."""

    def __init__(self, stager, slot, names, seqs, channel_major=False):
        self._stager, self._slot, self.names, self.seqs, self.channel_major = stager, slot, names, seqs, channel_major

    def release(self):
        """Return the arena to the ring (call after the batch's host->device copy has completed, e.g. after .result())."""
        if self._slot is not None:
            self._stager._free.put(self._slot)
            self._slot = None


class FeatureStager:
    """Iterate over ``videos`` in batches of ``batch_videos`` with the features staged in pinned host memory.

    feature_dir / videos / transpose: as ``Dataset`` passes them to ``load_feature`` (utils/dataset.py:12-21).
    dtype: torch.float32 (the reference format) or torch.bfloat16 (input-staging option for bf16 compute mode).
    depth: number of arenas; at most ``depth - 1`` batches may be un-released while the next one is being filled.
    pin: pinned arenas (needs a CUDA runtime); False gives pageable arenas (CPU-only tests).
    sort_by_length: batch videos of similar length together (less padding in the packed forward); the per-video results
    are keyed by name, so the order does not matter to a metrics pass.
    device_transpose (with transpose=True, fp32 arenas): keep the channel-major (D, T) layout of the files in the arena -- a
    straight memcpy instead of a strided host transpose, 4-5x faster on the host -- and let ``net.submit(...,
    channel_major=True)`` transpose (and cast) on the device; ``batch.seqs[i]`` is then (D, T_i) and
    ``batch.channel_major`` is True.
    """

    def __init__(self, feature_dir, videos, transpose=False, batch_videos=16, dtype=torch.float32, workers=4, depth=3,
                 pin=True, sort_by_length=False, device_transpose=False):
        assert dtype in (torch.float32, torch.bfloat16) and depth >= 2 and batch_videos >= 1
        assert not device_transpose or (transpose and dtype == torch.float32), 'device_transpose: channel-major fp32 arenas'
        self.dir, self.transpose, self.dtype, self.pin = feature_dir, transpose, dtype, pin
        self.device_transpose = device_transpose
        self.videos = list(videos)
        self._shape = {}
        for v in self.videos:                      # header-only reads: (T, D) after the optional transpose
            a = np.load(os.path.join(feature_dir, v + '.npy'), mmap_mode='r')
            assert a.ndim == 2, f'{v}: feature array must be 2-D, got {a.shape}'
            self._shape[v] = tuple(a.shape[::-1] if transpose else a.shape)
        dims = {s[1] for s in self._shape.values()}
        assert len(dims) == 1, f'videos disagree on the feature dimension: {sorted(dims)}'
        self.dim = dims.pop()
        order = sorted(self.videos, key=lambda v: (self._shape[v][0], v)) if sort_by_length else self.videos
        self.batches = [order[i:i + batch_videos] for i in range(0, len(order), batch_videos)]
        self._cap = max(sum(self._shape[v][0] for v in b) for b in self.batches) if self.batches else 0
        self._arenas = [None] * depth
        self._free = queue.Queue()
        for i in range(depth):
            self._free.put(i)
        self._pool = ThreadPoolExecutor(max_workers=max(workers, 1))

    def __len__(self):
        return len(self.batches)

    def frames(self):
        return sum(s[0] for s in self._shape.values())

    def _arena(self, slot):
        if self._arenas[slot] is None:
            self._arenas[slot] = torch.empty((self._cap, self.dim), dtype=self.dtype, pin_memory=self.pin)
        return self._arenas[slot]

    def _fill(self, dst, video):
        src = np.load(os.path.join(self.dir, video + '.npy'), mmap_mode='r')
        if self.transpose and not self.device_transpose:
            src = src.T
        if self.dtype == torch.float32:
            np.copyto(dst.numpy(), src, casting='unsafe')          # transpose + cast in one pass over the mapped file
        else:
            dst.copy_(torch.from_numpy(np.asarray(src)))           # strided, converting copy (round-to-nearest-even)

    def _stage(self, names, slot):
        arena, seqs, row = self._arena(slot), [], 0
        for v in names:
            T = self._shape[v][0]
            blk = arena[row:row + T]
            seqs.append(blk.view(self.dim, T) if self.device_transpose else blk)      # same bytes, file layout
            row += T
        list(self._pool.map(self._fill, seqs, names))
        return Batch(self, slot, list(names), seqs, channel_major=self.device_transpose)

    def __iter__(self):
        """Batches in order; batch i+1 is staged by a helper thread while the consumer works on batch i."""
        if not self.batches:
            return
        nxt = queue.Queue(maxsize=1)

        def producer():
            try:
                for names in self.batches:
                    slot = self._free.get()                        # blocks until the consumer releases an arena
                    nxt.put(self._stage(names, slot))
                nxt.put(None)
            except BaseException as e:                             # surface loader errors in the consumer thread
                nxt.put(e)

        threading.Thread(target=producer, daemon=True).start()
        while True:
            item = nxt.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item

    def close(self):
        self._pool.shutdown(wait=False)


def run_sweep(net, stager, labels=None):
    """Pipelined inference over a stager: yields (names, save_list) per batch, in order.  Keeps two batches in flight
    (the copy of batch i+1 overlaps the kernels of batch i, ``net.submit``) and releases an arena when its batch is done."""
    pending = []
    for batch in stager:
        ys = None if labels is None else [labels[n] for n in batch.names]
        pending.append((batch, net.submit(batch.seqs, ys, channel_major=batch.channel_major)))
        if len(pending) == 2:
            b, h = pending.pop(0)
            res = h.result()
            b.release()
            yield b.names, res
    for b, h in pending:
        res = h.result()
        b.release()
        yield b.names, res
