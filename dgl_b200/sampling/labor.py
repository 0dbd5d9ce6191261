"""LABOR layer-neighbor sampling (Balin & Catalyurek 2023).

Parity: /root/reference/python/dgl/sampling/labor.py, src/array/cuda/labor_sampling.cu,
src/array/cpu/labor_pick.h. LABOR-0: every *source* vertex t draws one uniform variate r_t shared by all
seeds; edge (t -> s) is kept iff r_t <= fanout / deg(s). Overlapping neighbourhoods therefore sample the
same vertices, shrinking the number of unique inputs. Importance sampling iterations refine per-vertex
probabilities pi_t so the expected number of sampled vertices is minimised.
"""
from __future__ import annotations

import torch

from .. import random as dgl_random
from ..base import EID, DGLError
from ..heterograph import _to_tensor
from .neighbor import _build_result


def _labor_rel(ug, seeds, fanout, prob, importance_sampling, random_seed, seed2_contribution=0.0):
    indptr, indices, eids = ug.csc()
    dev = seeds.device
    if (dev.type == "cuda" and prob is None and importance_sampling == 0 and not seed2_contribution
            and indptr.is_cuda and seeds.dtype == indptr.dtype):
        # LABOR-0: per-source variate r_t <= fanout / deg(s); count + fill kernels (csrc/cuda/graph_kernels.cu)
        from .._graph_kernels import labor0
        src, pos, eid, counts = labor0(indptr, indices, eids, seeds.contiguous(), fanout, random_seed)
        w = 1.0 / counts[pos.long()].clamp(min=1).float()     # Hajek-normalised: equal inclusion prob inside a seed
        return src, seeds[pos.long()], eid, w
    if dev.type == "cuda" and indptr.is_cuda and seeds.dtype == indptr.dtype and fanout >= 0:
        # LABOR-i / weighted / continuous-seed variants: importance iterations, normalisers and the Bernoulli pass are
        # device kernels over the CSC (csrc/cuda/labor.cu); one host read per call
        from .._graph_kernels import labor_sample
        p = None if prob is None else prob.to(dev).contiguous()
        src, pos, eid, w = labor_sample(indptr, indices, eids, seeds.contiguous(), fanout, p, importance_sampling,
                                        random_seed, int(random_seed) + 1, float(seed2_contribution), num_src=ug.num_src,
                                        cache=ug._degs)
        return src, seeds[pos.long()], eid, w
    indptr_d, indices_d = indptr.to(dev), indices.to(dev)
    s = indptr_d[seeds.long()].long()
    deg = indptr_d[seeds.long() + 1].long() - s
    total = int(deg.sum())
    pos = torch.repeat_interleave(s, deg) + (torch.arange(total, device=dev) - torch.repeat_interleave(torch.cumsum(deg, 0) - deg, deg))
    owner = torch.repeat_interleave(torch.arange(seeds.shape[0], device=dev), deg)
    nbr = indices_d[pos]
    e = eids.to(dev)[pos] if eids is not None else pos.to(ug.idtype)
    if fanout < 0 or total == 0:
        w = torch.ones(total, device=dev)
        return nbr, seeds[owner], e, w
    if prob is not None:
        a = prob.to(dev).float()[e.long()]
    else:
        a = torch.ones(total, device=dev)
    # per-seed normaliser so that sum_t min(1, c_s * pi_t * a_ts) == fanout
    uniq, inv = torch.unique(nbr, return_inverse=True)
    pi = torch.ones(uniq.shape[0], device=dev)
    k = torch.minimum(deg, torch.full_like(deg, fanout)).float()

    def solve_c(pi_edge):
        # binary search per seed for c_s
        lo = torch.zeros(seeds.shape[0], device=dev)
        hi = torch.full((seeds.shape[0],), 1.0, device=dev)
        # grow hi until enough mass
        for _ in range(40):
            mass = torch.zeros(seeds.shape[0], device=dev).index_add_(0, owner, torch.clamp(hi[owner] * pi_edge * a, max=1.0))
            need = mass < k - 1e-6
            if not bool(need.any()):
                break
            hi = torch.where(need, hi * 2, hi)
        for _ in range(30):
            mid = (lo + hi) / 2
            mass = torch.zeros(seeds.shape[0], device=dev).index_add_(0, owner, torch.clamp(mid[owner] * pi_edge * a, max=1.0))
            ge = mass >= k
            hi = torch.where(ge, mid, hi)
            lo = torch.where(ge, lo, mid)
        return hi

    iters = importance_sampling if importance_sampling >= 0 else 10
    c = solve_c(pi[inv])
    for _ in range(iters):
        # pi_t <- max_s c_s * a_ts * pi_t (fixed point iteration of LABOR-i)
        cand = c[owner] * pi[inv] * a
        newpi = torch.zeros_like(pi).scatter_reduce(0, inv, cand, reduce="amax", include_self=False)
        pi = newpi.clamp(min=1e-12)
        c = solve_c(pi[inv])
    p_edge = torch.clamp(c[owner] * pi[inv] * a, max=1.0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(random_seed) & 0x7FFFFFFF)
    r = torch.rand(uniq.shape[0], device=dev, generator=gen)
    if seed2_contribution > 0:
        gen2 = torch.Generator(device=dev)
        gen2.manual_seed((int(random_seed) + 1) & 0x7FFFFFFF)
        r2 = torch.rand(uniq.shape[0], device=dev, generator=gen2)
        # continuous interpolation between two seeds (graphbolt continuous_seed.h idea)
        import math
        t = seed2_contribution
        n1 = torch.erfinv(2 * r - 1) * math.sqrt(2)
        n2 = torch.erfinv(2 * r2 - 1) * math.sqrt(2)
        z = math.cos(math.pi * t / 2) * n1 + math.sin(math.pi * t / 2) * n2
        r = 0.5 * (1 + torch.erf(z / math.sqrt(2)))
    keep = r[inv] <= p_edge
    w = 1.0 / p_edge[keep]
    # normalise importance weights per seed (Hajek estimator)
    ow = owner[keep]
    sums = torch.zeros(seeds.shape[0], device=dev).index_add_(0, ow, w)
    w = w / sums[ow].clamp(min=1e-12)
    return nbr[keep], seeds[ow], e[keep], w


def sample_labors(g, nodes, fanout, edge_dir="in", prob=None, importance_sampling=0, random_seed=None,
                  seed2_contribution=0, copy_ndata=True, copy_edata=True, exclude_edges=None, output_device=None):
    """LABOR sampling; returns (subgraph, list of edge importance weights per etype)."""
    if edge_dir != "in":
        raise DGLError("sample_labors supports edge_dir='in' only")
    if not isinstance(nodes, dict):
        nodes = {g.ntypes[0]: nodes}
    fan = fanout if isinstance(fanout, dict) else {c: fanout for c in g.canonical_etypes}
    fan = {g.to_canonical_etype(k): int(v) for k, v in fan.items()}
    if random_seed is None:
        random_seed = dgl_random.next_seed()
    dev = None
    per, weights = [], []
    for i, c in enumerate(g.canonical_etypes):
        seeds = nodes.get(c[2])
        if seeds is None:
            e = torch.empty(0, dtype=g.idtype, device=g.device)
            per.append((e, e.clone(), e.clone())); weights.append(torch.empty(0, device=g.device))
            continue
        seeds = _to_tensor(seeds, g.idtype, seeds.device if torch.is_tensor(seeds) and seeds.is_cuda else g.device)
        dev = seeds.device
        p = g._edge_frames[i][prob] if prob is not None else None
        src, dst, eid, w = _labor_rel(g._graphs[i], seeds, fan[c], p, importance_sampling, random_seed,
                                      seed2_contribution)
        per.append((src, dst, eid)); weights.append(w)
    ret = _build_result(g, per, copy_ndata, copy_edata, dev if dev is not None else g.device)
    if output_device is not None:
        ret = ret.to(output_device)
        weights = [w.to(output_device) for w in weights]
    return ret, weights
