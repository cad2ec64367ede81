"""Data-parallel training step on 2 GPUs (NCCL): the all-reduced gradient equals the single-process batch gradient."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def _build(g, dev):
    from fact_clip_b200 import config as C
    from fact_clip_b200.loss import MatchCriterion
    from fact_clip_b200.models.blocks import FACT_CLIP
    from fact_clip_b200.utils.synth import make_text_embeddings
    cfg = C.tiny(**g['tiny_kwargs'])
    cfg.Loss.match, cfg.Loss.nullw, cfg.Loss.sw, cfg.Loss.pc = 'o2o', 0.1, 0.5, 0.2
    cfg.CLIP.projection_dropout = 0.0
    net = FACT_CLIP(cfg, g['in_dim'], g['n_classes'], make_text_embeddings(g['n_classes']))
    net.load_state_dict(g['state_dict'], strict=False)
    net.compute_mode = 'fp32'
    net = net.to(dev).train()
    net.mcriterion = MatchCriterion(cfg, g['n_classes'], [])
    return net


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from fact_clip_b200.parallel import GradAllReducer
    g = load_golden('tiny_m_iuU_clip_trained')
    net = _build(g, dev)
    red = GradAllReducer()
    net.grad_ready_hook = red.on_bucket
    v = g['videos'][rank]
    loss, _ = net([v['x'].to(dev)], [v['label'].to(dev)], compute_loss=True)
    loss.backward()
    nbytes = red.finish()
    torch.cuda.synchronize()
    if rank == 0:
        q.put(({n: p.grad.cpu() for n, p in net.named_parameters()}, nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduced_gradient_equals_batch_gradient():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    g = load_golden('tiny_m_iuU_clip_trained')
    net = _build(g, 'cuda:0')
    vids = g['videos'][:2]
    loss, _ = net([v['x'].to('cuda:0') for v in vids], [v['label'].to('cuda:0') for v in vids], compute_loss=True)
    loss.backward()
    ref = {n: p.grad.cpu() for n, p in net.named_parameters()}
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, nbytes = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert nbytes >= 4 * sum(p.numel() for p in net.parameters())
    gn = torch.sqrt(sum((r.double() ** 2).sum() for r in ref.values()))
    for n, r in ref.items():
        err = float((got[n].double() - r.double()).norm())
        assert err <= 1e-5 * max(float(r.double().norm()), 1e-3 * float(gn)), (n, err)
