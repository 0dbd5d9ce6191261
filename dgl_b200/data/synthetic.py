"""Synthetic graph generators shaped like the reference's benchmark datasets (no network access):
RMAT power-law graphs and uniform random graphs with the node/edge/feature counts cited in BASELINE.md."""
from __future__ import annotations

import torch

from ..convert import graph as make_graph

SHAPES = {
    # name: (num_nodes, num_undirected_or_directed_edges, feat_dim, num_classes, symmetrize)
    "cora": (2708, 10556, 1433, 7, False),
    "reddit": (232965, 114615892, 602, 41, False),
    "ogbn-products": (2449029, 61859140, 100, 47, True),
    "ogbn-arxiv": (169343, 1166243, 128, 40, False),
    "ogbn-papers100M": (111059956, 1615685872, 128, 172, False),
}


def rmat_edges(num_nodes, num_edges, a=0.57, b=0.19, c=0.19, seed=0, device="cpu", chunk=1 << 24):
    """R-MAT edge list (Chakrabarti et al.) with quadrant probabilities (a, b, c, 1-a-b-c); ids are
    scrambled with a fixed permutation-like hash so high-degree vertices are not clustered at 0."""
    scale = max(1, (num_nodes - 1).bit_length())
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    srcs, dsts = [], []
    done = 0
    while done < num_edges:
        m = min(chunk, num_edges - done)
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(m, device=device, generator=gen)
            sbit = (r >= a + b).long()
            dbit = ((r >= a) & (r < a + b) | (r >= a + b + c)).long()
            src = (src << 1) | sbit
            dst = (dst << 1) | dbit
        # scramble + fold into [0, num_nodes)
        src = (src * 0x9E3779B1 + 0x7F4A7C15) % (1 << scale)
        dst = (dst * 0x9E3779B1 + 0x7F4A7C15) % (1 << scale)
        src = src % num_nodes
        dst = dst % num_nodes
        srcs.append(src)
        dsts.append(dst)
        done += m
    return torch.cat(srcs), torch.cat(dsts)


def rmat_graph(num_nodes, num_edges, symmetrize=False, seed=0, device="cpu", idtype=torch.int32):
    u, v = rmat_edges(num_nodes, num_edges, seed=seed, device=device)
    if symmetrize:
        u, v = torch.cat([u, v]), torch.cat([v, u])
    return make_graph((u.to(idtype), v.to(idtype)), num_nodes=num_nodes, idtype=idtype, device=device)


def shaped_dataset(name, device="cpu", seed=0, idtype=torch.int32, feat_dtype=torch.float32, scale=1.0,
                   train_frac=0.0803):
    """Random-init dataset with the node/edge/feature/class counts of ``name`` (optionally scaled down).
    Returns (graph, feat, labels, train_idx, num_classes)."""
    n, e, f, c, sym = SHAPES[name]
    n = max(64, int(n * scale))
    e = max(256, int(e * scale))
    g = rmat_graph(n, e, symmetrize=sym, seed=seed, device=device, idtype=idtype)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    feat = torch.randn(n, f, device=device, generator=gen).to(feat_dtype)
    labels = torch.randint(0, c, (n,), device=device, generator=gen)
    ntrain = max(1, int(n * train_frac))
    train_idx = torch.randperm(n, device=device, generator=gen)[:ntrain]
    return g, feat, labels, train_idx, c


def __getattr__(name):
    # the explanation benchmarks of the reference's synthetic.py (BA-Shapes & co) are generated in extra.py
    if name in ("BAShapeDataset", "BACommunityDataset", "TreeCycleDataset", "TreeGridDataset", "BA2MotifDataset"):
        from . import extra
        return getattr(extra, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
