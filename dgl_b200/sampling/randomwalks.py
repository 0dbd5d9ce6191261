"""Random walks (parity: /root/reference/python/dgl/sampling/randomwalks.py:random_walk :18,
pack_traces :246; node2vec_randomwalk.py; native src/graph/sampling/randomwalks/*).
Walk steps are batched: every step advances all walkers with one vectorised CSR lookup, which maps to
a handful of bandwidth-bound kernels on the GPU (the reference runs thread-per-walk kernels)."""
from __future__ import annotations

import torch

from .. import random as dgl_random
from .._native import loader
from ..base import DGLError, id_code
from ..heterograph import _to_tensor

__all__ = ["random_walk", "pack_traces", "node2vec_random_walk"]


def _step(ug, cur, prob, gen_dev):
    """advance walkers at ``cur`` (−1 = terminated) along one relation; returns (next, eid)."""
    indptr, indices, eids = ug.csr()
    alive = cur >= 0
    c = cur.clamp(min=0).long()
    s = indptr[c].long()
    deg = indptr[c + 1].long() - s
    alive = alive & (deg > 0)
    if prob is None:
        r = torch.rand(cur.shape[0], device=gen_dev)
        off = (r * deg.clamp(min=1).to(r.dtype)).long().clamp(max=(deg - 1).clamp(min=0))
    else:
        # weighted choice: inverse-CDF inside each row via cumsum over the CSR-ordered weights
        w = prob[eids.long()] if eids is not None else prob
        w = w.to(torch.float64)
        cs = torch.cumsum(w, 0)
        base = torch.where(s > 0, cs[(s - 1).clamp(min=0)], torch.zeros_like(cs[:1]).expand_as(s))
        tot = cs[(s + deg - 1).clamp(min=0)] - base
        alive = alive & (tot > 0)
        target = base + torch.rand(cur.shape[0], device=gen_dev, dtype=torch.float64) * tot
        pos = torch.searchsorted(cs, target, right=True)
        off = (pos - s).clamp(min=0)
        off = torch.minimum(off, (deg - 1).clamp(min=0))
    p = (s + off).clamp(max=max(indices.shape[0] - 1, 0))
    nxt = torch.where(alive, indices[p] if indices.numel() else cur, torch.full_like(cur, -1))
    e = (eids[p] if eids is not None else p.to(cur.dtype)) if indices.numel() else cur
    e = torch.where(alive, e, torch.full_like(cur, -1))
    return nxt, e


def random_walk(g, nodes, *, metapath=None, length=None, prob=None, restart_prob=None, return_eids=False):
    """Random walks from ``nodes``; returns (traces [n, len+1], types [len+1]) (+ eids)."""
    n_etypes = len(g.canonical_etypes)
    if metapath is None:
        if n_etypes > 1 or len(g.ntypes) > 1:
            raise DGLError("metapath not specified and the graph is not homogeneous.")
        if length is None:
            raise ValueError("Please specify either the metapath or the random walk length.")
        metapath = [0] * length
    else:
        metapath = [g.get_etype_id(e) for e in metapath]
    idt = g.idtype
    dev = g.device
    nodes = _to_tensor(nodes, idt, dev)
    if dev.type == "cuda" and len(set(metapath)) <= 16 and len(metapath) > 0:
        # ONE kernel for the whole walk: thread per walker, relation table in kernel parameters
        from .._graph_kernels import random_walk as _rw
        uniq = sorted(set(metapath))
        rels = []
        for etid in uniq:
            indptr, indices, eids = g._graphs[etid].csr()
            rels.append((indptr, indices, eids, g._edge_frames[etid][prob] if prob is not None else None))
        step_rel = torch.tensor([uniq.index(e) for e in metapath], dtype=torch.int32, device=dev)
        rp = None
        if restart_prob is not None:
            rp = (restart_prob if torch.is_tensor(restart_prob) else torch.full((len(metapath),), float(restart_prob)))
            rp = rp.to(device=dev, dtype=torch.float32).contiguous()
        traces, eids_t = _rw(rels, step_rel, rp, nodes.contiguous(), len(metapath), dgl_random.next_seed(), return_eids)
        types = torch.tensor([g._src_ntid(metapath[0])] + [g._dst_ntid(e) for e in metapath], dtype=idt, device=dev)
        return (traces, eids_t, types) if return_eids else (traces, types)
    traces = [nodes]
    eids_out = []
    cur = nodes.clone()
    if restart_prob is not None and not torch.is_tensor(restart_prob):
        restart_prob = torch.full((len(metapath),), float(restart_prob), device=dev)
    for i, etid in enumerate(metapath):
        p = g._edge_frames[etid][prob] if prob is not None else None
        if restart_prob is not None:
            stop = torch.rand(cur.shape[0], device=dev) < restart_prob[i].to(dev)
            cur = torch.where(stop, torch.full_like(cur, -1), cur)
        cur, e = _step(g._graphs[etid], cur, p, dev)
        traces.append(cur)
        eids_out.append(e)
    traces = torch.stack(traces, 1)
    types = [g._src_ntid(metapath[0])] + [g._dst_ntid(e) for e in metapath] if metapath else [0]
    types = torch.tensor(types, dtype=idt, device=dev)
    if return_eids:
        return traces, torch.stack(eids_out, 1) if eids_out else torch.empty((nodes.shape[0], 0), dtype=idt), types
    return traces, types


def pack_traces(traces, types):
    """Concatenate valid (non -1) trace entries (reference randomwalks.py:246)."""
    mask = traces >= 0
    lengths = mask.sum(1)
    concat_vids = traces[mask]
    concat_types = types.to(traces.device).unsqueeze(0).expand_as(traces)[mask]
    offsets = torch.cumsum(lengths, 0) - lengths
    return concat_vids, concat_types, lengths, offsets


def node2vec_random_walk(g, nodes, p, q, walk_length, prob=None, return_eids=False):
    """Second-order biased walks (node2vec). Rejection sampling, thread per walker: native CPU kernel or the
    sm_100a kernel (csrc/cuda/graph_kernels.cu) depending on the graph's device."""
    if not g.is_homogeneous:
        raise DGLError("node2vec_random_walk requires a homogeneous graph")
    if prob is not None:
        raise DGLError("edge-weighted node2vec walks are not supported yet")
    idt = g.idtype
    out_dev = g.device
    ug = g._graphs[0]
    if out_dev.type == "cuda":
        from .._graph_kernels import node2vec_walk as _n2v
        indptr, indices, eids = ug.csr()
        seeds = _to_tensor(nodes, idt, out_dev).contiguous()
        traces, te = _n2v(indptr, indices, eids, seeds, walk_length, p, q, dgl_random.next_seed(), return_eids)
        return (traces, te) if return_eids else traces
    indptr, indices, eids = [None if t is None else t.cpu().contiguous() for t in ug.csr()]
    seeds = _to_tensor(nodes, idt, torch.device("cpu")).contiguous()
    n = seeds.shape[0]
    traces = torch.empty((n, walk_length + 1), dtype=idt)
    te = torch.empty((n, walk_length), dtype=idt)
    loader.cpu_lib().fn("dglb200_cpu_random_walk", "ippppll" + "ddd" + "lpp")(
        id_code(idt), indptr.data_ptr(), indices.data_ptr(), 0 if eids is None else eids.data_ptr(), seeds.data_ptr(),
        n, walk_length, 0.0, float(p), float(q), dgl_random.next_seed(), traces.data_ptr(), te.data_ptr())
    traces = traces.to(out_dev)
    if return_eids:
        return traces, te.to(out_dev)
    return traces
