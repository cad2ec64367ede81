"""Host-side logic of the multi-GPU sweep on CPU: world_size-2 gloo processes shard the video list, "predict"
and gather variable-length predictions on rank 0 (no GPU involved)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fact_clip_b200.parallel import GradAllReducer, gather_predictions, shard_by_length, shard_range


def fake_pred(i, T):
    return (np.arange(T, dtype=np.int64) * (i + 3)) % 17


def _worker(rank, world, port, lengths, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = shard_by_length(lengths, world)[rank]
    preds = [fake_pred(i, lengths[i]) for i in mine]
    merged = gather_predictions(mine, preds)
    # timing protocol of bench.py: max over ranks of the per-rank elapsed time
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((sorted(merged), all(np.array_equal(merged[i], fake_pred(i, lengths[i])) for i in merged), float(t)))
    else:
        assert merged is None
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_shard_and_gather_world2():
    lengths = [4096, 37, 1, 2000, 333, 4096, 5, 900, 901]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids, ok, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids == list(range(len(lengths))) and ok and tmax == 11.0


def test_shard_helpers():
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_range(3, 3, 4) == (3, 3)
    lengths = [5, 100, 7, 60, 50, 1]
    parts = shard_by_length(lengths, 2)
    assert sorted(sum(parts, [])) == list(range(6))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 10
    single = gather_predictions([4], [np.array([1, 2, 3])])          # world size 1: no process group needed
    assert list(single) == [4] and np.array_equal(single[4], [1, 2, 3])


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    red = GradAllReducer()
    g = torch.Generator().manual_seed(100 + rank)
    sizes = [1000, 37, 4096]                      # sections arrive in backward order, one flat buffer each
    flats = [torch.randn(n, generator=g) for n in sizes]
    for k, f in enumerate(flats):
        red.on_bucket(len(sizes) - 1 - k, f)
    nbytes = red.finish()
    if rank == 0:
        q.put(([f.clone() for f in flats], nbytes))
    dist.destroy_process_group()


def test_grad_allreduce_world2_equals_batch_mean():
    """The bucketed asynchronous all-reduce averages every section's flat gradient buffer in place: what rank 0 holds
    afterwards is the mean of the two ranks' gradients (= the gradient of the mean loss over both ranks' videos)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flats, nbytes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sizes = [1000, 37, 4096]
    assert nbytes == 4 * sum(sizes)
    gens = [torch.Generator().manual_seed(100 + r) for r in range(2)]
    for k, n in enumerate(sizes):
        want = (torch.randn(n, generator=gens[0]) + torch.randn(n, generator=gens[1])) / 2
        assert torch.allclose(flats[k], want, atol=1e-7)


def test_grad_reducer_single_process_is_a_noop():
    red = GradAllReducer()
    f = torch.arange(8.0)
    red.on_bucket(0, f)
    assert red.finish() == 32 and torch.equal(f, torch.arange(8.0))
