"""Multi-GPU plumbing for inference sweeps: videos are independent, so each rank owns a
contiguous shard and there is NO data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is
used only to gather the variable-length per-video predictions on rank 0 for the metrics pass
(fact_clip/utils/evaluate.py:96,230 in the reference needs whole per-video arrays) and for timing barriers."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_length(lengths, world):
    """Assign videos to ranks so that frames per rank are balanced (longest-first greedy); returns a list of
    index lists, each sorted by length so that batches have similar slot sizes."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    load, out = [0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += lengths[i]
    return [sorted(ix, key=lambda i: (lengths[i], i)) for ix in out]


def gather_predictions(local_ids, local_preds, group=None, dst=0):
    """Gather {video id -> int64 prediction array} on rank `dst`.  Uses padded tensor collectives (int32 payload)
    rather than pickling.  Returns the merged dict on dst, None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return {i: np.asarray(p, dtype=np.int64) for i, p in zip(local_ids, local_preds)}
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    lens = torch.tensor([len(p) for p in local_preds], dtype=torch.int64)
    meta = torch.tensor([len(local_ids), int(lens.sum()) if len(lens) else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_n, max_tot = max(int(m[0]) for m in metas), max(int(m[1]) for m in metas)
    head = torch.full((2, max(max_n, 1)), -1, dtype=torch.int64, device=dev)
    if len(local_ids):
        head[0, :len(local_ids)] = torch.tensor(list(local_ids), dtype=torch.int64)
        head[1, :len(local_ids)] = lens
    body = torch.zeros(max(max_tot, 1), dtype=torch.int32, device=dev)
    if len(local_preds):
        flat = np.concatenate([np.asarray(p) for p in local_preds]).astype(np.int32)
        body[:len(flat)] = torch.from_numpy(flat).to(dev)
    heads = [torch.zeros_like(head) for _ in range(world)]
    bodies = [torch.zeros_like(body) for _ in range(world)]
    dist.all_gather(heads, head, group=group)
    dist.all_gather(bodies, body, group=group)
    if rank != dst:
        return None
    out = {}
    for h, b in zip(heads, bodies):
        h, b, off = h.cpu(), b.cpu().numpy(), 0
        for vid, n in zip(h[0].tolist(), h[1].tolist()):
            if vid < 0:
                continue
            out[vid] = b[off:off + n].astype(np.int64)
            off += n
    return out


class GradAllReducer:
    """Data-parallel gradient averaging for the training step (the reference trains on one GPU, the training
    bench of BASELINE config 5 is data parallel: one video per GPU per step).

    ``net.grad_ready_hook = reducer.on_bucket`` -- the training engine calls it with each section's flat gradient buffer as
    soon as the backward pass has left that section (CLIP head first, input block last); the all-reduce (SUM, in place, NCCL
    over NVLink on GPUs, gloo in CPU tests) is launched asynchronously and overlaps the rest of the backward pass.
    ``finish()`` waits for the collectives and divides by the world size: the loss is a mean over videos (blocks.py:130, 914),
    so with equal videos per rank the averaged gradient IS the single-process batch gradient.  ``clip_grad_norm_`` and the
    optimizer step (scripts/train.py:266-268) must come after ``finish()``."""

    def __init__(self, group=None):
        self.group, self.handles, self.buckets = group, [], []
        self.bytes = 0

    def on_bucket(self, k, flat, names=None):
        self.buckets.append(flat)
        self.bytes += flat.numel() * flat.element_size()
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            self.handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        for h in self.handles:
            h.wait()
        if world > 1 and self.buckets:
            torch._foreach_mul_(self.buckets, 1.0 / world)
        n = self.bytes
        self.handles, self.buckets, self.bytes = [], [], 0
        return n


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated:
    pinned pages are placed on the allocating thread's node, and a feature buffer that sits behind the inter-socket link
    caps the host->device copies of an 8-GPU sweep well below what the two sockets' memory can feed.  Reads sysfs only
    (PCI address from the CUDA device properties); returns a small report dict, never raises.  FACTK_NUMA_BIND=0 disables."""
    import os
    rep = {'bound': False}
    if os.environ.get('FACTK_NUMA_BIND', '1') == '0':
        rep['why'] = 'disabled'
        return rep
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f'{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as f:
            node = int(f.read().strip())
        rep.update(pci=bdf, node=node)
        if node < 0:
            rep['why'] = 'no NUMA information for the device'
            return rep
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = _parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            rep['why'] = 'no allowed CPU on the node'
            return rep
        os.sched_setaffinity(0, cpus)
        rep.update(bound=True, cpus=len(cpus))
    except Exception as e:          # topology files missing in a container, old torch without pci ids, ...
        rep['why'] = f'{type(e).__name__}: {e}'
    return rep
