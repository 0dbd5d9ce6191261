"""Neighbor sampling (``dgl.sampling.sample_neighbors`` family).

Parity: /root/reference/python/dgl/sampling/neighbor.py (sample_etype_neighbors :69,
sample_neighbors :222, sample_neighbors_fused :399, sample_neighbors_biased :690,
select_topk :880); native: csrc/cpu/sampling.cc and csrc/cuda/sampling.cu. The graph may live
on the CPU, on the GPU, or in host-pinned memory with GPU seeds (UVA sampling).
"""
from __future__ import annotations

import torch

from .. import random as dgl_random
from .._native import loader
from ..base import EID, NID, DGLError, id_code
from ..frame import Frame
from ..graph_index import UnitGraph
from ..heterograph import DGLGraph, _to_tensor

__all__ = ["sample_neighbors", "sample_etype_neighbors", "sample_neighbors_fused", "sample_neighbors_biased",
           "select_topk", "sample_rel"]


_TOPK_WDTYPE = {torch.float32: 0, torch.float64: 1, torch.int64: 2}


def _p(t):
    return 0 if t is None else t.data_ptr()


def sample_rel(ug, seeds, fanout, edge_dir="in", prob=None, replace=False, seed=None, return_counts=False):
    """Sample one relation. Returns (src, dst, eid) tensors on the seeds' device."""
    if edge_dir == "in":
        indptr, indices, eids = ug.csc()
    else:
        indptr, indices, eids = ug.csr()
    idt = ug.idtype
    seeds = seeds.to(idt).contiguous()
    n = seeds.shape[0]
    dev = seeds.device
    graph_dev = indptr.device
    uva = dev.type == "cuda" and graph_dev.type == "cpu"
    if uva and not indptr.is_pinned():
        raise DGLError("UVA sampling needs a pinned graph: call g.pin_memory_() first")
    if dev.type == "cpu" and graph_dev.type == "cuda":
        raise DGLError("seed nodes are on CPU but the graph is on GPU")
    if seed is None:
        seed = dgl_random.next_seed()
    if prob is not None:
        prob = prob.to(torch.float32).contiguous()
        if uva and not prob.is_cuda:
            prob = prob.to(dev)
    if n == 0:
        e = torch.empty(0, dtype=idt, device=dev)
        if return_counts:
            return e, e.clone(), e.clone(), torch.empty(0, dtype=torch.int64, device=dev)
        return e, e.clone(), e.clone()
    counts = torch.empty(n, dtype=torch.int64, device=dev)
    if dev.type == "cuda":
        lib = loader.cuda_lib()
        st = loader.stream_ptr()
        lib.fn("dglb200_cuda_sample_count", "ippipllippp")(
            id_code(idt), _p(indptr), _p(eids), 1 if uva else 0, _p(seeds), n, int(fanout), int(bool(replace)),
            _p(prob), _p(counts), st)
    else:
        lib = loader.cpu_lib()
        lib.fn("dglb200_cpu_sample_count", "ipppllipp")(
            id_code(idt), _p(indptr), _p(eids), _p(seeds), n, int(fanout), int(bool(replace)), _p(prob), _p(counts))
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=offsets[1:])
    total = int(offsets[-1].item())
    src = torch.empty(total, dtype=idt, device=dev)
    dst = torch.empty(total, dtype=idt, device=dev)
    eid = torch.empty(total, dtype=idt, device=dev)
    if total > 0:
        if dev.type == "cuda":
            lib.fn("dglb200_cuda_sample_fill", "ipppipllipplpppp")(
                id_code(idt), _p(indptr), _p(indices), _p(eids), 1 if uva else 0, _p(seeds), n, int(fanout),
                int(bool(replace)), _p(prob), _p(offsets), seed, _p(src), _p(dst), _p(eid), st)
        else:
            lib.fn("dglb200_cpu_sample_fill", "ippppllipplppp")(
                id_code(idt), _p(indptr), _p(indices), _p(eids), _p(seeds), n, int(fanout), int(bool(replace)),
                _p(prob), _p(offsets), seed, _p(src), _p(dst), _p(eid))
    if edge_dir == "out":
        # rows were sources: native fills (neighbor, seed) => (dst, src)
        src, dst = dst, src
    if return_counts:
        return src, dst, eid, counts
    return src, dst, eid


def _exclude(src, dst, eid, exclude):
    if exclude is None or exclude.numel() == 0:
        return src, dst, eid
    ex = torch.sort(exclude.to(eid.device).to(eid.dtype))[0]
    pos = torch.searchsorted(ex, eid).clamp(max=ex.numel() - 1)
    keep = ex[pos] != eid
    return src[keep], dst[keep], eid[keep]


def _build_result(g, per_etype, copy_ndata, copy_edata, out_dev):
    graphs = []
    for i, (src, dst, eid) in enumerate(per_etype):
        ug = g._graphs[i]
        graphs.append(UnitGraph.from_coo(ug.num_src, ug.num_dst, src, dst))
    ret = DGLGraph.__new__(DGLGraph)
    ret._init(graphs, None if g._srctypes is not None else list(g._ntypes), list(g._canonical_etypes),
              list(g._num_nodes), None, None, g._srctypes, g._dsttypes)
    if copy_ndata and g.device == out_dev:
        ret._node_frames = [f.clone() for f in g._node_frames]
    for i, (src, dst, eid) in enumerate(per_etype):
        fr = Frame(num_rows=int(eid.shape[0]))
        if copy_edata and len(g._edge_frames[i]) > 0 and g.device == out_dev:
            sub = g._edge_frames[i].subframe(eid)
            fr._columns = dict(sub._columns)
        fr[EID] = eid
        ret._edge_frames[i] = fr
    return ret


def sample_neighbors(g, nodes, fanout, edge_dir="in", prob=None, replace=False, copy_ndata=True, copy_edata=True,
                     _dist_training=False, exclude_edges=None, output_device=None):
    """For every seed pick up to ``fanout`` in-edges (or out-edges) per edge type, uniformly or
    proportionally to the edge feature ``prob``; ``fanout=-1`` keeps all edges. Returns a graph on
    the same node-id space whose edges carry the original ids in ``edata[dgl.EID]``."""
    if edge_dir not in ("in", "out"):
        raise DGLError("edge_dir must be 'in' or 'out'")
    if not hasattr(g, "_graphs") and hasattr(g, "gpb"):  # DistGraph: per-partition RPC sampling
        return g.sample_neighbors(nodes, fanout, edge_dir=edge_dir, prob=prob, replace=replace)
    if g.is_block:
        raise DGLError("Sampling from a block is not supported")
    if not isinstance(nodes, dict):
        if len(g.ntypes) > 1:
            raise DGLError("Must specify node type when the graph is not homogeneous.")
        nodes = {g.ntypes[0]: nodes}
    dev = None
    nodes_t = {}
    for k, v in nodes.items():
        t = v if torch.is_tensor(v) else torch.as_tensor(v)
        if dev is None:
            dev = t.device if t.is_cuda or g.device.type == "cpu" else g.device
        nodes_t[k] = t.to(device=dev, dtype=g.idtype).reshape(-1)
    if dev is None:
        dev = g.device
    if isinstance(fanout, dict):
        fan = {g.to_canonical_etype(k): int(v) for k, v in fanout.items()}
        if len(fan) != len(g.canonical_etypes):
            raise DGLError("Fan-out must be specified for each edge type if a dict is provided.")
    else:
        fan = {c: int(fanout) for c in g.canonical_etypes}
    if isinstance(exclude_edges, dict):
        excl = {g.to_canonical_etype(k): _to_tensor(v, g.idtype, dev) for k, v in exclude_edges.items()}
    elif exclude_edges is not None:
        excl = {g.canonical_etypes[0]: _to_tensor(exclude_edges, g.idtype, dev)}
    else:
        excl = {}
    per_etype = []
    for i, c in enumerate(g.canonical_etypes):
        seed_type = c[2] if edge_dir == "in" else c[0]
        seeds = nodes_t.get(seed_type)
        k = fan[c]
        if seeds is None or seeds.numel() == 0 or k == 0:
            e = torch.empty(0, dtype=g.idtype, device=dev)
            per_etype.append((e, e.clone(), e.clone()))
            continue
        p = None
        if prob is not None:
            if prob not in g._edge_frames[i]:
                raise DGLError(f"probability feature '{prob}' not found on edge type {c}")
            p = g._edge_frames[i][prob]
            if p.dtype == torch.bool or not p.is_floating_point():
                p = p.float()
        src, dst, eid = sample_rel(g._graphs[i], seeds, k, edge_dir, p, replace)
        src, dst, eid = _exclude(src, dst, eid, excl.get(c))
        per_etype.append((src, dst, eid))
    ret = _build_result(g, per_etype, copy_ndata, copy_edata, dev)
    return ret if output_device is None else ret.to(output_device)


def _native_pick(symbol, sig, idt, n, head, tail_after_offsets):
    """Two-phase driver of the host row-wise pickers in csrc/cpu/rowwise_pick.cc: the count pass (offsets = NULL) sizes the
    outputs, the fill pass writes (neighbour, seed, edge id). ``head`` are the arguments before ``offsets``,
    ``tail_after_offsets`` those between ``offsets`` and ``counts``."""
    lib = loader.cpu_lib()
    fn = lib.fn(symbol, sig)
    counts = torch.empty(n, dtype=torch.int64)
    fn(*head, 0, *tail_after_offsets, _p(counts), 0, 0, 0)
    offsets = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(counts, 0, out=offsets[1:])
    total = int(offsets[-1])
    nb, sd, eid = (torch.empty(total, dtype=idt) for _ in range(3))
    if total:
        fn(*head, _p(offsets), *tail_after_offsets, _p(counts), _p(nb), _p(sd), _p(eid))
    return nb, sd, eid


def sample_etype_neighbors(g, nodes, etype_offset, fanout, edge_dir="in", prob=None, replace=False, copy_ndata=True,
                           copy_edata=True, etype_sorted=False, _dist_training=False, output_device=None):
    """Per-edge-type fan-outs on a homogenised graph whose edges are grouped by type:
    ``etype_offset[t]:etype_offset[t+1]`` is the edge-id range of type t (reference neighbor.py:69)."""
    if not g.is_homogeneous:
        raise DGLError("sample_etype_neighbors expects a homogenised graph")
    idt = g.idtype
    nodes = _to_tensor(nodes, idt, nodes.device if torch.is_tensor(nodes) and nodes.is_cuda else g.device)
    fan = fanout if torch.is_tensor(fanout) else torch.as_tensor(fanout)
    ntype_edges = len(etype_offset) - 1
    if fan.numel() == 1:
        fan = fan.repeat(ntype_edges)
    if g.device.type == "cpu" and nodes.device.type == "cpu":
        # host path: one native pass buckets every row's neighbours by edge type and draws each bucket's fan-out
        ug = g._graphs[0]
        indptr, indices, eids_ = ug.csc() if edge_dir == "in" else ug.csr()
        off = torch.as_tensor(etype_offset, dtype=torch.int64).contiguous()
        fan64 = fan.to(torch.int64).contiguous()
        p = None if prob is None else g.edata[prob].to(torch.float32).contiguous()
        seeds = nodes.contiguous()
        nb, sd, eid = _native_pick(
            "dglb200_cpu_etype_pick", "ippppLpLpippLpppp".replace("L", "l"), idt, seeds.shape[0],
            (id_code(idt), _p(indptr), _p(indices), _p(eids_), _p(seeds), seeds.shape[0], _p(off), ntype_edges, _p(fan64),
             int(bool(replace)), _p(p)), (dgl_random.next_seed(),))
        src, dst = (nb, sd) if edge_dir == "in" else (sd, nb)
        ret = _build_result(g, [(src, dst, eid)], copy_ndata, copy_edata, nodes.device)
        return ret if output_device is None else ret.to(output_device)
    srcs, dsts, eids = [], [], []
    u, v = g._graphs[0].coo()
    for t in range(ntype_edges):
        lo, hi = int(etype_offset[t]), int(etype_offset[t + 1])
        if hi <= lo or int(fan[t]) == 0:
            continue
        sub = UnitGraph.from_coo(g.num_nodes(), g.num_nodes(), u[lo:hi], v[lo:hi])
        p = None
        if prob is not None:
            p = g.edata[prob][lo:hi].float()
        s, d, e = sample_rel(sub, nodes, int(fan[t]), edge_dir, p, replace)
        srcs.append(s); dsts.append(d); eids.append(e + lo)
    if srcs:
        src, dst, eid = torch.cat(srcs), torch.cat(dsts), torch.cat(eids)
    else:
        src = dst = eid = torch.empty(0, dtype=idt, device=nodes.device)
    ret = _build_result(g, [(src, dst, eid)], copy_ndata, copy_edata, nodes.device)
    return ret if output_device is None else ret.to(output_device)


def sample_neighbors_fused(g, nodes, fanout, edge_dir="in", prob=None, replace=False, copy_ndata=True,
                           copy_edata=True, exclude_edges=None, mapping=None):
    """Sample and compact in one call: returns (block, induced_src_nodes) (reference neighbor.py:399)."""
    from ..transforms import to_block
    fr = sample_neighbors(g, nodes, fanout, edge_dir=edge_dir, prob=prob, replace=replace, copy_ndata=copy_ndata,
                          copy_edata=copy_edata, exclude_edges=exclude_edges)
    block = to_block(fr, nodes if edge_dir == "in" else None)
    induced = {nt: block.srcnodes[nt].data[NID] for nt in block.srctypes}
    if len(induced) == 1:
        induced = next(iter(induced.values()))
    return block, induced


def sample_neighbors_biased(g, nodes, fanout, bias, edge_dir="in", tag_offset_name="_TAG_OFFSET", replace=False,
                            copy_ndata=True, copy_edata=True, output_device=None):
    """Sampling biased by the tag of the neighbour node; neighbours must be tag-sorted with
    ``sort_csr_by_tag``/``sort_csc_by_tag`` (reference neighbor.py:690)."""
    if not g.is_homogeneous and len(g.etypes) > 1:
        raise DGLError("Only homogeneous or bipartite graphs are supported")
    ug = g._graphs[0]
    indptr, indices, eids = ug.csc() if edge_dir == "in" else ug.csr()
    off = (g.dstdata if edge_dir == "in" else g.srcdata)[tag_offset_name]
    if indptr.device.type == "cpu":
        # host path: native two-level draw over the tag ranges, no per-edge probability array is materialised
        seeds = _to_tensor(nodes, g.idtype, g.device).contiguous()
        off64 = off.to(torch.int64).contiguous()
        b64 = torch.as_tensor(bias, dtype=torch.float64).contiguous()
        if off64.shape[1] != b64.shape[0] + 1:
            raise DGLError(f"bias has {b64.shape[0]} entries but the tag offsets describe {off64.shape[1] - 1} tags")
        nb, sd, eid = _native_pick(
            "dglb200_cpu_biased_pick", "ippppLpLpLipLpppp".replace("L", "l"), g.idtype, seeds.shape[0],
            (id_code(g.idtype), _p(indptr), _p(indices), _p(eids), _p(seeds), seeds.shape[0], _p(off64), b64.shape[0],
             _p(b64), int(fanout), int(bool(replace))), (dgl_random.next_seed(),))
        src, dst = (nb, sd) if edge_dir == "in" else (sd, nb)
        ret = _build_result(g, [(src, dst, eid)], copy_ndata, copy_edata, seeds.device)
        return ret if output_device is None else ret.to(output_device)
    bias = torch.as_tensor(bias, dtype=torch.float32, device=indptr.device)
    deg = (indptr[1:] - indptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(indptr.shape[0] - 1, device=indptr.device), deg)
    pos_in_row = torch.arange(indices.shape[0], device=indptr.device) - indptr[:-1].long()[rows]
    tag = (pos_in_row.unsqueeze(1) >= off[rows][:, 1:].long()).sum(1)
    p_pos = bias[tag.clamp(max=bias.shape[0] - 1)]
    e = eids.long() if eids is not None else torch.arange(indices.shape[0], device=indptr.device)
    prob = torch.empty(indices.shape[0], dtype=torch.float32, device=indptr.device)
    prob[e] = p_pos
    seeds = _to_tensor(nodes, g.idtype, g.device)
    src, dst, eid = sample_rel(ug, seeds, fanout, edge_dir, prob, replace)
    ret = _build_result(g, [(src, dst, eid)], copy_ndata, copy_edata, seeds.device)
    return ret if output_device is None else ret.to(output_device)


def select_topk(g, k, weight, nodes=None, edge_dir="in", ascending=False, copy_ndata=True, copy_edata=True,
                output_device=None):
    """Keep the k in-edges (or out-edges) with the largest/smallest ``weight`` per node
    (reference neighbor.py:880)."""
    if nodes is None:
        nodes = {nt: torch.arange(g.num_nodes(nt), dtype=g.idtype, device=g.device) for nt in g.ntypes}
    elif not isinstance(nodes, dict):
        nodes = {g.ntypes[0]: nodes}
    kd = k if isinstance(k, dict) else {c: k for c in g.canonical_etypes}
    kd = {g.to_canonical_etype(c): v for c, v in kd.items()}
    per_etype = []
    for i, c in enumerate(g.canonical_etypes):
        seed_type = c[2] if edge_dir == "in" else c[0]
        ug = g._graphs[i]
        if seed_type not in nodes or kd[c] == 0:
            e = torch.empty(0, dtype=g.idtype, device=g.device)
            per_etype.append((e, e.clone(), e.clone()))
            continue
        seeds = _to_tensor(nodes[seed_type], g.idtype, g.device)
        indptr, indices, eids = ug.csc() if edge_dir == "in" else ug.csr()
        wt = g._edge_frames[i][weight]
        if indptr.device.type == "cpu" and wt.dim() == 1 and wt.dtype in _TOPK_WDTYPE:
            # host path: per-row partial sort in native code (csrc/cpu/rowwise_pick.cc), stable on ties
            seeds = seeds.contiguous()
            deg = indptr[seeds.long() + 1].long() - indptr[seeds.long()].long()
            cnt = deg if kd[c] < 0 else deg.clamp(max=int(kd[c]))
            offsets = torch.zeros(seeds.shape[0] + 1, dtype=torch.int64)
            torch.cumsum(cnt, 0, out=offsets[1:])
            total = int(offsets[-1])
            nb, sd, ee = (torch.empty(total, dtype=g.idtype) for _ in range(3))
            if total:
                wt = wt.contiguous()
                loader.cpu_lib().fn("dglb200_cpu_topk_pick", "iipppplplipppp")(
                    id_code(g.idtype), _TOPK_WDTYPE[wt.dtype], _p(indptr), _p(indices), _p(eids), _p(seeds), seeds.shape[0],
                    _p(wt), int(kd[c]), int(bool(ascending)), _p(offsets), _p(nb), _p(sd), _p(ee))
            per_etype.append((nb, sd, ee) if edge_dir == "in" else (sd, nb, ee))
            continue
        s = indptr[seeds.long()].long()
        cnt = indptr[seeds.long() + 1].long() - s
        pos = torch.repeat_interleave(s, cnt) + (torch.arange(int(cnt.sum()), device=s.device)
                                                 - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
        owner = torch.repeat_interleave(torch.arange(seeds.shape[0], device=s.device), cnt)
        e = eids[pos].long() if eids is not None else pos
        w = g._edge_frames[i][weight][e].double()
        kk = kd[c]
        if kk < 0:
            sel = torch.arange(pos.shape[0], device=s.device)
        else:
            key = w if ascending else -w
            order = torch.sort(key, stable=True)[1]
            order = order[torch.sort(owner[order], stable=True)[1]]
            starts = torch.cumsum(cnt, 0) - cnt
            rank = torch.arange(order.shape[0], device=s.device) - starts[owner[order]]
            sel = order[rank < kk]
        nb = indices[pos[sel]]
        sd = seeds[owner[sel]]
        ee = e[sel].to(g.idtype)
        per_etype.append((nb, sd, ee) if edge_dir == "in" else (sd, nb, ee))
    ret = _build_result(g, per_etype, copy_ndata, copy_edata, g.device)
    return ret if output_device is None else ret.to(output_device)
