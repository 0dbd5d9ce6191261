"""Approximate k-nearest-neighbour graphs by NN-descent (Dong, Charikar, Li 2011), tensorised.

Reference: /root/reference/src/graph/transform/cuda/knn.cu:261-766 (one thread block per point walking candidate lists
with per-point locks) and src/graph/transform/cpu/knn.cc NNDescent. Here a round is a handful of dense tensor operations
over ALL points at once, so the same code runs on the CPU and on the GPU and the distance evaluations are batched:

  1. candidates of p = (sampled) members of B(u) for every u in B(p), B = forward list + (capped) reverse list,
  2. distances point -> candidates in row chunks (bounded memory),
  3. merge with the current list: duplicates and self are masked, the k smallest survive,
  4. stop when fewer than ``delta * N * k`` entries changed.

Segments (several point sets in one tensor) are handled by drawing the initial lists inside each segment: candidates are
reached through existing neighbours only, so they never leave the segment.
"""
from __future__ import annotations

import torch


def _sq_dist(x, rows, cand):
    """||x[rows] - x[cand]||^2 for cand [R, M] -> [R, M] (fp32)."""
    a = x[rows].float()                                  # [R, D]
    b = x[cand].float()                                  # [R, M, D]
    return (a * a).sum(1, keepdim=True) + (b * b).sum(2) - 2.0 * torch.einsum("rd,rmd->rm", a, b)


def _reverse_lists(nbr, width):
    """up to ``width`` reverse neighbours per point ([N, width], padded with the point itself)."""
    n, k = nbr.shape
    dev = nbr.device
    src = torch.arange(n, device=dev).repeat_interleave(k)
    dst = nbr.reshape(-1)
    order = torch.argsort(dst, stable=True)
    dst_s, src_s = dst[order], src[order]
    start = torch.searchsorted(dst_s, torch.arange(n, device=dev))
    rank = torch.arange(dst_s.numel(), device=dev) - start[dst_s]
    keep = rank < width
    out = torch.arange(n, device=dev).unsqueeze(1).repeat(1, width)
    out[dst_s[keep], rank[keep]] = src_s[keep]
    return out


def nn_descent(x, offsets, k, max_iters=12, sample=None, delta=0.001, chunk_elems=1 << 26, generator=None):
    """[N, k] ids of (approximately) the k nearest points of the same segment, self included like the exact kernels
    (a point's nearest neighbour is itself). ``offsets``: [S + 1] segment boundaries (host tensor / list)."""
    n, d = x.shape
    dev = x.device
    offsets = [int(o) for o in offsets]
    sample = max(2, min(k, 6)) if sample is None else int(sample)
    # ---- initial lists: k distinct random points of the own segment (self first)
    nbr = torch.empty((n, k), dtype=torch.long, device=dev)
    for s in range(len(offsets) - 1):
        lo, hi = offsets[s], offsets[s + 1]
        m = hi - lo
        if m == 0:
            continue
        base = torch.arange(lo, hi, device=dev)
        if k - 1 >= m:                                           # tiny segment: every other point is a neighbour
            ring = (torch.arange(m, device=dev).unsqueeze(1) + torch.arange(k, device=dev).unsqueeze(0)) % m
            nbr[lo:hi] = lo + ring
            continue
        # distinct random offsets 1..m-1 per row: a random start plus k-1 distinct strides of a random permutation prefix
        shift = torch.rand(m, m - 1, device=dev, generator=generator).argsort(1)[:, : k - 1] + 1 if m <= 2048 else \
            torch.stack([torch.randperm(m - 1, device=dev, generator=generator)[: k - 1] + 1 for _ in range(1)]).expand(m, -1)
        if m > 2048:                                             # decorrelate the rows of the shared permutation
            shift = (shift + torch.randint(0, m - 1, (m, 1), device=dev, generator=generator)) % (m - 1) + 1
        nbr[lo:hi, 0] = base
        nbr[lo:hi, 1:] = lo + (torch.arange(m, device=dev).unsqueeze(1) + shift) % m
    rows_per_chunk = max(1, chunk_elems // max(1, (2 * k + 2 * k * sample) * d))
    dist = torch.empty((n, k), dtype=torch.float32, device=dev)
    for r0 in range(0, n, rows_per_chunk):
        r = torch.arange(r0, min(n, r0 + rows_per_chunk), device=dev)
        dist[r] = _sq_dist(x, r, nbr[r])
    for _ in range(max_iters):
        # ---- candidates: `sample` random neighbours of every neighbour + up to k reverse neighbours
        # local join over the UNDIRECTED neighbourhood B(p) = forward list + (capped) reverse list: the candidates of p are
        # `sample` random members of B(u) for every u in B(p), plus B(p)'s reverse half itself
        rev = _reverse_lists(nbr, k)
        both = torch.cat([nbr, rev], dim=1)                                  # [N, 2k]
        pick = torch.randint(0, 2 * k, (n, 2 * k, sample), device=dev, generator=generator)
        two_hop = torch.gather(both[both], 2, pick).reshape(n, 2 * k * sample)
        cand = torch.cat([two_hop, rev], dim=1)
        changed = 0
        for r0 in range(0, n, rows_per_chunk):
            r = torch.arange(r0, min(n, r0 + rows_per_chunk), device=dev)
            ids = torch.cat([nbr[r], cand[r]], dim=1)
            dd = torch.cat([dist[r], _sq_dist(x, r, cand[r])], dim=1)
            # mask duplicates: after sorting by id, equal neighbours are adjacent; keep the first of every run
            order = torch.argsort(ids, dim=1, stable=True)
            ids_s, dd_s = torch.gather(ids, 1, order), torch.gather(dd, 1, order)
            dup = torch.zeros_like(ids_s, dtype=torch.bool)
            dup[:, 1:] = ids_s[:, 1:] == ids_s[:, :-1]
            dd_s = dd_s.masked_fill(dup, float("inf"))
            best = torch.topk(dd_s, k, dim=1, largest=False)
            new_ids = torch.gather(ids_s, 1, best.indices)
            changed += int((new_ids.sort(1).values != nbr[r].sort(1).values).sum())
            nbr[r], dist[r] = new_ids, best.values
        if changed <= delta * n * k:
            break
    return nbr
