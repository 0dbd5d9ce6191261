"""Time-respecting neighbour sampling on a CSC graph, homogeneous or typed.

For a seed ``v`` with timestamp ``t_v`` only in-edges whose source (node timestamp) and / or own (edge timestamp) time
lies in ``[t_v - window, t_v)`` are eligible (``window`` optional). Reference behaviour:
/root/reference/graphbolt/src/fused_csc_sampling_graph.cc:482-770 (TemporalSampleNeighbors - per-seed loops with a
temporal picker) and python/dgl/graphbolt/impl/fused_csc_sampling_graph.py:817-1080.

This implementation is data-parallel end to end, so the same code runs on the CPU and on the GPU without a host
round-trip per seed: candidate edges of all seeds are expanded into one flat list (owner, CSC position), masked by the
temporal predicate, given one random key each and cut to the ``fanout`` smallest keys per owner with two stable sorts.
Keys are ``Exp(1) / p`` (Efraimidis-Spirakis), i.e. uniform without replacement when ``p`` is constant and weighted
without replacement otherwise; the layer variant derives the key from a counter-based hash of (seed, source id) so that
a source vertex receives the SAME variate for every seed that can reach it (the LABOR coupling, which shrinks the union
of sampled sources).
"""
from __future__ import annotations

import torch

_M64 = (1 << 63) - 1


def _hash_uniform(ids, seed):
    """counter-based U(0,1) variate per id (splitmix-style avalanche on int64 lanes, sign bit kept clear)."""
    x = (ids.to(torch.int64) + int(seed) * 0x9E3779B97F4A7C15 % _M64) & _M64
    for mult, shift in ((0x3F58476D1CE4E5B9, 30), (0x14D049BB133111EB, 27)):
        x = ((x ^ (x >> shift)) * (mult % _M64)) & _M64
    x = x ^ (x >> 31)
    return ((x >> 10) & ((1 << 53) - 1)).to(torch.float64) / float(1 << 53)


def expand_candidates(indptr, seeds):
    """(owner, pos): for every in-edge of every seed its seed ordinal and CSC position."""
    seeds = seeds.long()
    start = indptr[seeds].long()
    deg = indptr[seeds + 1].long() - start
    total = int(deg.sum())
    owner = torch.repeat_interleave(torch.arange(seeds.shape[0], device=seeds.device), deg, output_size=total)
    first = torch.cumsum(deg, 0) - deg
    pos = start[owner] + (torch.arange(total, device=seeds.device) - first[owner])
    return owner, pos


def temporal_mask(owner, pos, indices, seed_ts, window, node_ts, edge_ts):
    ok = torch.ones(pos.shape[0], dtype=torch.bool, device=pos.device)
    upper = seed_ts[owner]
    lower = None if window is None else upper - window[owner]
    for stamps in ((None if node_ts is None else node_ts[indices[pos].long()]), (None if edge_ts is None else edge_ts[pos])):
        if stamps is None:
            continue
        ok &= stamps < upper
        if lower is not None:
            ok &= stamps >= lower
    return ok


def pick_per_owner(owner, pos, keep, num_owners, fanout, replace, weights=None, variates=None):
    """Choose up to ``fanout`` eligible candidates per owner. Returns (selected candidate indices grouped by owner, counts).
    ``variates``: optional U(0,1) per candidate replacing the private random stream (layer sampling)."""
    cand = keep.nonzero().squeeze(1)
    if weights is not None:
        cand = cand[weights[cand] > 0]
    own = owner[cand]
    if fanout < 0 or cand.numel() == 0:
        return cand, torch.bincount(own, minlength=num_owners)
    if replace:
        # fanout independent draws per owner that has at least one candidate: inverse-CDF inside the owner's segment
        w = torch.ones(cand.shape[0], dtype=torch.float64, device=cand.device) if weights is None else weights[cand].double()
        csum = torch.cumsum(w, 0)
        cnt = torch.bincount(own, minlength=num_owners)
        seg_end = torch.cumsum(cnt, 0)
        seg_start = seg_end - cnt
        base = torch.where(seg_start > 0, csum[(seg_start - 1).clamp(min=0)], torch.zeros_like(csum[:1]).expand(num_owners))
        total = csum[(seg_end - 1).clamp(min=0)] - base
        live = (cnt > 0).nonzero().squeeze(1)
        u = torch.rand(live.shape[0], fanout, dtype=torch.float64, device=cand.device)
        target = base[live].unsqueeze(1) + u * total[live].unsqueeze(1)
        hit = torch.searchsorted(csum, target.reshape(-1), right=True)
        hit = torch.minimum(hit, (seg_end[live] - 1).repeat_interleave(fanout)).clamp(min=0)
        hit = torch.maximum(hit, seg_start[live].repeat_interleave(fanout))
        out_cnt = torch.zeros(num_owners, dtype=torch.int64, device=cand.device)
        out_cnt[live] = fanout
        return cand[hit], out_cnt
    u = torch.rand(cand.shape[0], dtype=torch.float64, device=cand.device) if variates is None else variates[cand]
    key = -torch.log1p(-u.clamp(max=1 - 1e-16))                      # Exp(1)
    if weights is not None:
        key = key / weights[cand].double()
    by_key = torch.argsort(key)
    by_owner = torch.argsort(own[by_key], stable=True)
    order = by_key[by_owner]                                          # grouped by owner, ascending key inside a group
    cnt = torch.bincount(own, minlength=num_owners)
    first = torch.cumsum(cnt, 0) - cnt
    rank = torch.arange(order.shape[0], device=cand.device) - first[own[order]]
    chosen = order[rank < fanout]
    return cand[chosen], torch.clamp(cnt, max=fanout)


def temporal_sample(indptr, indices, seeds, seed_ts, fanout, replace=False, window=None, node_ts=None, edge_ts=None,
                    probs=None, edge_filter=None, layer_seed=None):
    """One relation. Returns (source ids [global], per-seed counts, CSC positions) grouped by seed order.
    ``edge_filter``: optional bool per CSC position (e.g. "edge is of this type")."""
    owner, pos = expand_candidates(indptr, seeds)
    keep = temporal_mask(owner, pos, indices, seed_ts, window, node_ts, edge_ts)
    if edge_filter is not None:
        keep &= edge_filter[pos]
    weights = None if probs is None else probs[pos].double()
    variates = None if layer_seed is None else _hash_uniform(indices[pos], layer_seed)
    sel, cnt = pick_per_owner(owner, pos, keep, seeds.shape[0], int(fanout), replace, weights, variates)
    sel_pos = pos[sel]
    return indices[sel_pos], cnt, sel_pos
