"""Deterministic synthetic inputs: segment-structured per-frame features.

x_t = proto[class_t] + sigma * N(0, I), proto ~ N(0, I) of shape (C, D); ``nseg`` uniformly drawn
boundaries; labels are the segment classes.  No dataset or checkpoint is read (there is none).
"""
import torch


def make_video(T, in_dim, n_classes, seed, nseg=8, sigma=0.5, proto_seed=7):
    g = torch.Generator().manual_seed(int(seed))
    proto = torch.randn(n_classes, in_dim, generator=torch.Generator().manual_seed(proto_seed))
    nseg = max(1, min(nseg, T))
    cuts = torch.sort(torch.randperm(max(T - 1, 1), generator=g)[:nseg - 1] + 1).values if T > 1 else torch.zeros(0, dtype=torch.long)
    bounds = torch.cat([torch.zeros(1, dtype=torch.long), cuts, torch.tensor([T])])
    classes = torch.randint(0, n_classes, (nseg,), generator=g)
    label = torch.zeros(T, dtype=torch.long)
    for s in range(len(bounds) - 1):
        label[bounds[s]:bounds[s + 1]] = classes[s]
    x = proto[label] + sigma * torch.randn(T, in_dim, generator=g)
    return x.contiguous(), label


def make_batch(lengths, in_dim, n_classes, base_seed=0, nseg=8, sigma=0.5):
    vids = [make_video(T, in_dim, n_classes, base_seed + i, nseg, sigma) for i, T in enumerate(lengths)]
    return [v[0] for v in vids], [v[1] for v in vids]


def make_text_embeddings(n_classes, dim=512, seed=2):
    t = torch.randn(n_classes, dim, generator=torch.Generator().manual_seed(seed))
    return torch.nn.functional.normalize(t, dim=-1)
