"""Global uniform negative sampling (parity: /root/reference/python/dgl/sampling/negative.py,
src/array/cuda/negative_sampling.cu): rejection-sample node pairs that are not edges."""
from __future__ import annotations

import torch

from ..base import DGLError


def global_uniform_negative_sampling(g, num_samples, exclude_self_loops=True, replace=False, etype=None,
                                     redundancy=None):
    c = g.to_canonical_etype(etype)
    ug = g._graphs[g.get_etype_id(c)]
    ns, nd = ug.num_src, ug.num_dst
    if exclude_self_loops and c[0] != c[2]:
        exclude_self_loops = False
    if g.device.type == "cuda":
        return _neg_cuda(g, ug, ns, nd, int(num_samples), exclude_self_loops, replace, redundancy)
    row, col = ug.coo()
    dev = row.device
    key = torch.sort(row.long() * nd + col.long())[0]
    redundancy = 1.3 if redundancy is None else 1 + redundancy
    want = int(num_samples)
    got_u, got_v, seen = [], [], None
    have = 0
    for _ in range(8):
        m = int((want - have) * redundancy) + 16
        u = torch.randint(0, ns, (m,), device=dev)
        v = torch.randint(0, nd, (m,), device=dev)
        q = u * nd + v
        ok = torch.ones(m, dtype=torch.bool, device=dev)
        if key.numel():
            pos = torch.searchsorted(key, q).clamp(max=key.numel() - 1)
            ok &= key[pos] != q
        if exclude_self_loops:
            ok &= u != v
        u, v, q = u[ok], v[ok], q[ok]
        if not replace:
            allq = q if seen is None else torch.cat([seen, q])
            uq, first = torch.unique(allq, return_inverse=False, return_counts=False, sorted=True), None
            # keep first occurrences of new keys only
            if seen is not None:
                isnew = ~torch.isin(q, seen)
                u, v, q = u[isnew], v[isnew], q[isnew]
            _, idx = torch.unique(q, return_inverse=True)
            firstpos = torch.full((int(idx.max()) + 1 if idx.numel() else 0,), q.numel(), device=dev, dtype=torch.long)
            if idx.numel():
                firstpos.scatter_reduce_(0, idx, torch.arange(q.numel(), device=dev), reduce="amin")
                keep = torch.zeros(q.numel(), dtype=torch.bool, device=dev)
                keep[firstpos] = True
                u, v, q = u[keep], v[keep], q[keep]
            seen = q if seen is None else torch.cat([seen, q])
        got_u.append(u); got_v.append(v)
        have += u.numel()
        if have >= want:
            break
    u = torch.cat(got_u)[:want].to(g.idtype)
    v = torch.cat(got_v)[:want].to(g.idtype)
    return u, v


def _neg_cuda(g, ug, ns, nd, want, exclude_self_loops, replace, redundancy):
    """Rejection sampling on the device (csrc/cuda/graph_kernels.cu neg_sample_kernel): one trial per thread, the
    edge test walks the CSR row of u; accepted pairs are compacted (and de-duplicated unless ``replace``)."""
    from .. import random as dgl_random
    from .._graph_kernels import neg_sample
    indptr, indices, _ = ug.csr()
    redundancy = 1.3 if redundancy is None else 1 + redundancy
    got_u, got_v, seen, have = [], [], None, 0
    for _ in range(8):
        trials = int((want - have) * redundancy) + 16
        u, v, ok = neg_sample(indptr, indices, False, ns, nd, exclude_self_loops, trials, dgl_random.next_seed())
        u, v = u[ok], v[ok]
        if not replace:
            q = u.long() * nd + v.long()
            if seen is not None:
                fresh = ~torch.isin(q, seen)
                u, v, q = u[fresh], v[fresh], q[fresh]
            uq, inv = torch.unique(q, return_inverse=True)
            first = torch.full((uq.shape[0],), q.shape[0], device=q.device, dtype=torch.long)
            first.scatter_reduce_(0, inv, torch.arange(q.shape[0], device=q.device), reduce="amin")
            first = torch.sort(first)[0]
            u, v, q = u[first], v[first], q[first]
            seen = q if seen is None else torch.cat([seen, q])
        got_u.append(u); got_v.append(v)
        have += u.shape[0]
        if have >= want:
            break
    return torch.cat(got_u)[:want].to(g.idtype), torch.cat(got_v)[:want].to(g.idtype)
