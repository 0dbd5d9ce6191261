"""Functional graph transforms.

Parity: /root/reference/python/dgl/transforms/functional.py (45 functions) and
transforms/to_block.py:28-249. Block construction on CUDA tensors uses the
order-preserving hash-compaction kernel (csrc/cuda/compact.cu); on CPU the native
ordered hash table (csrc/cpu/sampling.cc).
"""
from __future__ import annotations

import math
from collections import defaultdict

import numpy as np
import torch

from .. import function as fn
from .._native import loader
from ..base import EID, ETYPE, NID, NTYPE, DGLError, dgl_warning, id_code
from ..convert import graph as make_graph
from ..convert import heterograph as make_heterograph
from ..convert import to_heterogeneous, to_homogeneous
from ..frame import Frame
from ..graph_index import UnitGraph
from ..heterograph import DGLBlock, DGLGraph, _to_tensor

__all__ = [
    "reverse", "add_self_loop", "remove_self_loop", "add_reverse_edges", "to_bidirected", "to_simple", "to_block",
    "compact_graphs", "khop_adj", "khop_graph", "line_graph", "add_nodes", "add_edges", "remove_nodes",
    "remove_edges", "laplacian_lambda_max", "metapath_reachable_graph", "reorder_graph", "sort_csr_by_tag",
    "sort_csc_by_tag", "knn_graph", "segmented_knn_graph", "radius_graph", "random_walk_pe", "lap_pe",
    "laplacian_pe", "double_radius_node_labeling", "shortest_dist", "svd_pe", "to_half", "to_bfloat16", "to_float", "to_double",
    "adj_product_graph", "adj_sum_graph", "out_subgraph_like", "unique_and_compact", "norm_by_dst",
    "reverse_heterograph", "to_simple_graph", "to_bidirected_stale", "laplacian_lambda_max",
    "metis_partition", "metis_partition_assignment", "partition_graph_with_halo",
    "knn", "metis_perm", "rcmk_perm", "pairwise_squared_distance",
]


def _copy_frames(src_g, dst_g, copy_ndata=True, copy_edata=True):
    if copy_ndata:
        dst_g._node_frames = [f.clone() for f in src_g._node_frames]
    if copy_edata:
        dst_g._edge_frames = [f.clone() for f in src_g._edge_frames]


def reverse(g, copy_ndata=True, copy_edata=False, *, share_ndata=None, share_edata=None):
    """Reverse every edge; relation (A, r, B) becomes (B, r, A) (reference functional.py:reverse)."""
    if share_ndata is not None:
        copy_ndata = share_ndata
    if share_edata is not None:
        copy_edata = share_edata
    if g.is_block:
        raise DGLError("Reversing a block graph is not supported.")
    cets = [(d, e, s) for (s, e, d) in g.canonical_etypes]
    order = sorted(range(len(cets)), key=lambda i: cets[i])
    graphs = [g._graphs[i].reverse() for i in order]
    new = DGLGraph(graphs, g.ntypes, [cets[i] for i in order], list(g._num_nodes))
    if copy_ndata:
        new._node_frames = [f.clone() for f in g._node_frames]
    if copy_edata:
        new._edge_frames = [g._edge_frames[i].clone() for i in order]
    return new


reverse_heterograph = reverse


def add_self_loop(g, edge_feat_names=None, fill_data=1.0, etype=None):
    etid = g.get_etype_id(etype)
    s, _, d = g.canonical_etypes[etid]
    if s != d:
        raise DGLError(f"add_self_loop does not support unidirectional bipartite graphs: {etype}. "
                       "Please make sure the types of head node and tail node are identical.")
    new = g.clone()
    n = g.num_nodes(s)
    nodes = torch.arange(n, dtype=g.idtype, device=g.device)
    data = None
    if g.num_edges(etype) > 0 and isinstance(fill_data, str):
        # aggregate incoming edge features per node as the new edge features
        data = {}
        names = edge_feat_names if edge_feat_names is not None else list(g._edge_frames[etid].keys())
        reducer = {"sum": fn.sum, "mean": fn.mean, "max": fn.max, "min": fn.min}[fill_data]
        with g.local_scope():
            for k in names:
                g.update_all(fn.copy_e(k, "m"), reducer("m", "_agg"), etype=etype)
                data[k] = g.nodes[s].data["_agg"]
    elif not isinstance(fill_data, str):
        names = edge_feat_names if edge_feat_names is not None else list(g._edge_frames[etid].keys())
        data = {}
        for k in names:
            ref = g._edge_frames[etid][k]
            data[k] = torch.full((n,) + tuple(ref.shape[1:]), fill_data, dtype=ref.dtype, device=ref.device)
    new.add_edges(nodes, nodes, data=data, etype=etype)
    return new


def remove_self_loop(g, etype=None):
    etid = g.get_etype_id(etype)
    s, _, d = g.canonical_etypes[etid]
    if s != d:
        raise DGLError(f"remove_self_loop does not support unidirectional bipartite graphs: {etype}.")
    u, v = g.all_edges(etype=etype)
    eid = torch.nonzero(u == v).squeeze(1)
    new = g.clone()
    new.remove_edges(eid, etype=etype)
    return new


def add_reverse_edges(g, readonly=None, copy_ndata=True, copy_edata=False, ignore_bipartite=False, exclude_self=True):
    """Add a reverse edge for every edge (reference functional.py:add_reverse_edges)."""
    data = {}
    feats = {}
    for i, c in enumerate(g.canonical_etypes):
        s, _, d = c
        u, v = g._graphs[i].coo()
        if s != d:
            if not ignore_bipartite:
                raise DGLError("add_reverse_edges is not well defined for unidirectional bipartite graphs, but "
                               f"{c} is unidirectional bipartite")
            data[c] = (u, v)
            feats[c] = None
            continue
        if exclude_self:
            m = u != v
            ru, rv = v[m], u[m]
            sel = torch.nonzero(m).squeeze(1)
        else:
            ru, rv = v, u
            sel = torch.arange(u.shape[0], device=u.device)
        data[c] = (torch.cat([u, ru]), torch.cat([v, rv]))
        feats[c] = sel
    new = make_heterograph(data, {t: g.num_nodes(t) for t in g.ntypes}, idtype=g.idtype, device=g.device) \
        if not g.is_homogeneous else make_graph(data[g.canonical_etypes[0]], num_nodes=g.num_nodes(), idtype=g.idtype,
                                                device=g.device)
    if copy_ndata:
        new._node_frames = [f.clone() for f in g._node_frames]
    if copy_edata:
        for i, c in enumerate(g.canonical_etypes):
            j = new.get_etype_id(c if not g.is_homogeneous else None)
            sel = feats[c]
            for k in g._edge_frames[i].keys():
                val = g._edge_frames[i][k]
                new._edge_frames[j][k] = val if sel is None else torch.cat([val, val[sel]], 0)
    return new


def to_bidirected(g, copy_ndata=False, readonly=None):
    """Simple bidirected graph: add reverse edges and drop duplicates (reference to_bidirected)."""
    for c in g.canonical_etypes:
        if c[0] != c[2]:
            raise DGLError("to_bidirected is not well defined for unidirectional bipartite graphs, but "
                           f"{c} is unidirectional bipartite")
    g2 = add_reverse_edges(g, copy_ndata=copy_ndata, copy_edata=False, exclude_self=False)
    return to_simple(g2, return_counts=None, copy_ndata=copy_ndata, copy_edata=False)


to_bidirected_stale = to_bidirected


def to_simple(g, return_counts="count", writeback_mapping=False, copy_ndata=True, copy_edata=False,
              aggregator="arbitrary"):
    """Remove parallel edges; optionally store multiplicities (reference to_simple)."""
    if g.is_block:
        raise DGLError("Cannot convert a block graph to a simple graph.")
    graphs, counts, maps, firsts = [], [], [], []
    for i, ug in enumerate(g._graphs):
        row, col = ug.coo()
        if not row.is_cuda:
            # host: native two-pass counting sort on (row, col) with group ids, multiplicities and first members
            from ..ops import cpu_struct
            r, c, inverse, cnt, first = cpu_struct.coalesce(row, col, ug.num_src, ug.num_dst)
            graphs.append(UnitGraph.from_coo(ug.num_src, ug.num_dst, r, c))
            counts.append(cnt)
            maps.append(inverse.to(g.idtype))
            firsts.append(first)
            continue
        nd = max(ug.num_dst, 1)
        key = row.long() * nd + col.long()
        uniq, inverse, cnt = torch.unique(key, return_inverse=True, return_counts=True)
        r = (uniq // nd).to(g.idtype)
        c = (uniq % nd).to(g.idtype)
        graphs.append(UnitGraph.from_coo(ug.num_src, ug.num_dst, r, c))
        counts.append(cnt)
        maps.append(inverse.to(g.idtype))
        first = torch.full((uniq.shape[0],), key.shape[0], dtype=torch.int64, device=key.device)
        first.scatter_reduce_(0, inverse, torch.arange(key.shape[0], device=key.device), reduce="amin")
        firsts.append(first)
    new = DGLGraph(graphs, g.ntypes, g.canonical_etypes, list(g._num_nodes))
    if copy_ndata:
        new._node_frames = [f.clone() for f in g._node_frames]
    if copy_edata:
        for i in range(len(graphs)):
            for k in g._edge_frames[i].keys():
                val = g._edge_frames[i][k]
                if aggregator == "arbitrary":
                    new._edge_frames[i][k] = val[firsts[i]]
                else:
                    out = torch.zeros((graphs[i].num_edges(),) + tuple(val.shape[1:]), dtype=val.dtype, device=val.device)
                    idx = maps[i].long().view(-1, *([1] * (val.dim() - 1))).expand_as(val)
                    out.scatter_add_(0, idx, val)
                    if aggregator == "mean":
                        out = out / counts[i].view(-1, *([1] * (val.dim() - 1))).to(out.dtype)
                    new._edge_frames[i][k] = out
    if return_counts is not None:
        for i in range(len(graphs)):
            new._edge_frames[i][return_counts] = counts[i].to(g.idtype)
    if writeback_mapping:
        if g.is_homogeneous:
            return new, maps[0]
        return new, {c: m for c, m in zip(g.canonical_etypes, maps)}
    return new


to_simple_graph = to_simple


def unique_and_compact(dst_nodes, src_ids):
    """Order-preserving compaction: returns (unique_nodes, local_src_ids) where ``unique_nodes``
    starts with ``dst_nodes`` followed by new ids in first-occurrence order."""
    idt = src_ids.dtype
    if dst_nodes is None:
        dst_nodes = torch.empty(0, dtype=idt, device=src_ids.device)
    n_first, n = dst_nodes.shape[0], src_ids.shape[0]
    if src_ids.is_cuda:
        from ..ops.compact import cuda_unique_compact
        return cuda_unique_compact(dst_nodes, src_ids)
    out_unique = torch.empty(n_first + n, dtype=idt)
    out_local = torch.empty(n, dtype=idt)
    cnt = torch.zeros(1, dtype=torch.int64)
    loader.cpu_lib().fn("dglb200_cpu_unique_compact", "iplplppp")(
        id_code(idt), dst_nodes.contiguous().data_ptr(), n_first, src_ids.contiguous().data_ptr(), n,
        out_unique.data_ptr(), out_local.data_ptr(), cnt.data_ptr())
    return out_unique[: int(cnt[0])], out_local


def to_block(g, dst_nodes=None, include_dst_in_src=True, src_nodes=None):
    """Convert a graph (usually a sampled frontier) into a message-flow graph whose dst nodes are
    ``dst_nodes`` and whose src nodes are [dst_nodes ++ remaining sources] in first-occurrence order
    (reference transforms/to_block.py:28-249)."""
    idt, dev = g.idtype, g.device
    if dst_nodes is None:
        dst_nodes = {}
        for i, c in enumerate(g.canonical_etypes):
            _, col = g._graphs[i].coo()
            dst_nodes.setdefault(c[2], []).append(col)
        dst_nodes = {t: torch.unique(torch.cat(v)) for t, v in dst_nodes.items()}
    elif not isinstance(dst_nodes, dict):
        if len(g.ntypes) > 1:
            raise DGLError("Graph has more than one node type; please specify a dict for dst_nodes.")
        dst_nodes = {g.ntypes[0]: dst_nodes}
    dst_nodes = {t: _to_tensor(v, idt, dev) for t, v in dst_nodes.items()}
    for t in g.ntypes:
        dst_nodes.setdefault(t, torch.empty(0, dtype=idt, device=dev))
    if src_nodes is not None and not isinstance(src_nodes, dict):
        src_nodes = {g.ntypes[0]: src_nodes}
    # gather src ids per source type
    src_cat = defaultdict(list)
    for i, c in enumerate(g.canonical_etypes):
        row, _ = g._graphs[i].coo()
        src_cat[c[0]].append((i, row))
    src_unique, src_local = {}, {}
    for t in g.ntypes:
        lst = src_cat.get(t, [])
        ids = torch.cat([r for _, r in lst]) if lst else torch.empty(0, dtype=idt, device=dev)
        if src_nodes is not None:
            uniq = _to_tensor(src_nodes[t], idt, dev)
            from ..ops.compact import map_ids
            local = map_ids(uniq, ids)
        else:
            first = dst_nodes[t] if include_dst_in_src else torch.empty(0, dtype=idt, device=dev)
            uniq, local = unique_and_compact(first, ids)
        src_unique[t] = uniq
        off = 0
        for i, r in lst:
            src_local[i] = local[off: off + r.shape[0]]
            off += r.shape[0]
    # map dst ids
    from ..ops.compact import map_ids
    graphs = []
    for i, c in enumerate(g.canonical_etypes):
        _, col = g._graphs[i].coo()
        dloc = map_ids(dst_nodes[c[2]], col)
        if col.numel() and bool((dloc < 0).any()):
            raise DGLError("Some destination nodes of the edges are not in dst_nodes.")
        graphs.append(UnitGraph.from_coo(src_unique[c[0]].shape[0], dst_nodes[c[2]].shape[0], src_local[i], dloc))
    srctypes = list(g.ntypes)
    dsttypes = list(g.ntypes)
    num_nodes = [int(src_unique[t].shape[0]) for t in srctypes] + [int(dst_nodes[t].shape[0]) for t in dsttypes]
    block = DGLBlock(graphs, None, g.canonical_etypes, num_nodes, srctypes=srctypes, dsttypes=dsttypes)
    # features: induced ids + sliced parent features
    for j, t in enumerate(srctypes):
        fr = g._node_frames[g.get_ntype_id(t)].subframe(src_unique[t])
        fr[NID] = src_unique[t]
        block._node_frames[j] = fr
    for j, t in enumerate(dsttypes):
        fr = g._node_frames[g.get_ntype_id(t)].subframe(dst_nodes[t])
        fr[NID] = dst_nodes[t]
        block._node_frames[len(srctypes) + j] = fr
    for i in range(len(graphs)):
        fr = g._edge_frames[i].clone()
        if EID not in fr:
            fr[EID] = torch.arange(graphs[i].num_edges(), dtype=idt, device=dev)
        block._edge_frames[i] = fr
    return block


def compact_graphs(graphs, always_preserve=None, copy_ndata=True, copy_edata=True):
    """Drop isolated nodes consistently across graphs sharing a node space (reference compact_graphs)."""
    return_single = False
    if not isinstance(graphs, (list, tuple)):
        graphs = [graphs]
        return_single = True
    if len(graphs) == 0:
        return []
    g0 = graphs[0]
    idt, dev = g0.idtype, g0.device
    if always_preserve is None:
        always_preserve = {}
    elif not isinstance(always_preserve, dict):
        always_preserve = {g0.ntypes[0]: always_preserve}
    keep = {}
    for t in g0.ntypes:
        ids = [_to_tensor(always_preserve[t], idt, dev)] if t in always_preserve else []
        for g in graphs:
            for i, c in enumerate(g.canonical_etypes):
                row, col = g._graphs[i].coo()
                if c[0] == t:
                    ids.append(row)
                if c[2] == t:
                    ids.append(col)
        cat = torch.cat(ids) if ids else torch.empty(0, dtype=idt, device=dev)
        first = _to_tensor(always_preserve[t], idt, dev) if t in always_preserve else None
        uniq, _ = unique_and_compact(first, cat)
        keep[t] = uniq
    from ..ops.compact import map_ids
    outs = []
    for g in graphs:
        ugs = []
        for i, c in enumerate(g.canonical_etypes):
            row, col = g._graphs[i].coo()
            ugs.append(UnitGraph.from_coo(keep[c[0]].shape[0], keep[c[2]].shape[0], map_ids(keep[c[0]], row),
                                          map_ids(keep[c[2]], col)))
        new = DGLGraph(ugs, g.ntypes, g.canonical_etypes, [int(keep[t].shape[0]) for t in g.ntypes])
        for j, t in enumerate(g.ntypes):
            if copy_ndata:
                fr = g._node_frames[j].subframe(keep[t])
                new._node_frames[j] = fr
            new._node_frames[j][NID] = keep[t]
        if copy_edata:
            new._edge_frames = [f.clone() for f in g._edge_frames]
        outs.append(new)
    return outs[0] if return_single else outs


def add_nodes(g, num, data=None, ntype=None):
    g = g.clone()
    g.add_nodes(num, data=data, ntype=ntype)
    return g


def add_edges(g, u, v, data=None, etype=None):
    g = g.clone()
    g.add_edges(u, v, data=data, etype=etype)
    return g


def remove_edges(g, eids, etype=None, store_ids=False):
    g = g.clone()
    g.remove_edges(eids, etype=etype, store_ids=store_ids)
    return g


def remove_nodes(g, nids, ntype=None, store_ids=False):
    g = g.clone()
    g.remove_nodes(nids, ntype=ntype, store_ids=store_ids)
    return g


def _scipy_adj(g, transpose=False):
    import scipy.sparse as sp
    u, v = g.all_edges()
    u, v = u.cpu().numpy(), v.cpu().numpy()
    n = g.num_nodes()
    m = sp.coo_matrix((np.ones(len(u), dtype=np.float64), (u, v)), shape=(n, n)).tocsr()
    return m.T.tocsr() if transpose else m


def khop_adj(g, k):
    """k-hop adjacency as a dense tensor (A^T)^k (reference khop_adj)."""
    a = _scipy_adj(g, transpose=True)
    return torch.from_numpy((a ** k).todense()).float()


def khop_graph(g, k, copy_ndata=True):
    """Graph whose edges connect k-hop neighbours (multiplicity = #paths)."""
    a = _scipy_adj(g)
    ak = (a ** k).tocoo()
    rep = ak.data.astype(np.int64)
    row = np.repeat(ak.row, rep)
    col = np.repeat(ak.col, rep)
    new = make_graph((torch.from_numpy(row.astype(np.int64)), torch.from_numpy(col.astype(np.int64))),
                     num_nodes=g.num_nodes(), idtype=g.idtype, device=g.device)
    if copy_ndata:
        new._node_frames = [f.clone() for f in g._node_frames]
    return new


def line_graph(g, backtracking=True, shared=False):
    """Line graph: node i = edge i of g; edge (i, j) iff dst(i) == src(j) (reference line_graph)."""
    if not g.is_homogeneous:
        raise DGLError("only homogeneous graph is supported")
    u, v = g.all_edges()
    E = u.shape[0]
    indptr, _, eids = g._graphs[0].csr()
    if not u.is_cuda:
        from ..ops import cpu_struct
        li, lj = cpu_struct.line_graph(u.to(g.idtype), v.to(g.idtype), indptr, eids, backtracking)
        lg = make_graph((li, lj), num_nodes=E, idtype=g.idtype, device=g.device)
        if shared:
            lg._node_frames = [g._edge_frames[0].clone()]
        return lg
    # for each edge i (u_i -> v_i) connect to all out-edges j of v_i
    s = indptr[v.long()].long()
    cnt = indptr[v.long() + 1].long() - s
    li = torch.repeat_interleave(torch.arange(E, device=u.device), cnt)
    pos = torch.repeat_interleave(s, cnt) + (torch.arange(int(cnt.sum()), device=u.device)
                                             - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
    lj = eids[pos].long() if eids is not None else pos
    if not backtracking:
        m = ~((u[li] == v[lj]) & (v[li] == u[lj]))
        li, lj = li[m], lj[m]
    lg = make_graph((li.to(g.idtype), lj.to(g.idtype)), num_nodes=E, idtype=g.idtype, device=g.device)
    if shared:
        lg._node_frames = [g._edge_frames[0].clone()]
    return lg


def laplacian_lambda_max(g):
    """Largest eigenvalue of the normalised symmetric Laplacian of each graph in a batch."""
    from ..batch import unbatch
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    rst = []
    for g_i in unbatch(g):
        n = g_i.num_nodes()
        adj = _scipy_adj(g_i, transpose=True).astype(float)
        deg = np.asarray(g_i.in_degrees().cpu().numpy(), dtype=float).clip(1)
        norm = sp.diags(deg ** -0.5)
        lap = sp.eye(n) - norm @ adj @ norm
        rst.append(float(spl.eigs(lap, 1, which="LM", return_eigenvectors=False)[0].real))
    return rst


def metapath_reachable_graph(g, metapath):
    """Graph connecting nodes reachable via ``metapath`` (reference metapath_reachable_graph)."""
    import scipy.sparse as sp
    adj = 1
    for etype in metapath:
        c = g.to_canonical_etype(etype)
        u, v = g.all_edges(etype=c)
        m = sp.coo_matrix((np.ones(u.shape[0]), (u.cpu().numpy(), v.cpu().numpy())),
                          shape=(g.num_nodes(c[0]), g.num_nodes(c[2]))).tocsr()
        adj = adj * m if not isinstance(adj, int) else m
    adj = (adj != 0).tocoo()
    srctype = g.to_canonical_etype(metapath[0])[0]
    dsttype = g.to_canonical_etype(metapath[-1])[2]
    u = torch.from_numpy(adj.row.astype(np.int64))
    v = torch.from_numpy(adj.col.astype(np.int64))
    if srctype == dsttype:
        new = make_graph((u, v), num_nodes=g.num_nodes(srctype), idtype=g.idtype, device=g.device)
        new._node_frames = [g._node_frames[g.get_ntype_id(srctype)].clone()]
        new._ntypes[0] = srctype
        return DGLGraph(new._graphs, [srctype], [(srctype, "_E", srctype)], [g.num_nodes(srctype)], new._node_frames)
    new = make_heterograph({(srctype, "_E", dsttype): (u, v)}, {srctype: g.num_nodes(srctype), dsttype: g.num_nodes(dsttype)},
                           idtype=g.idtype, device=g.device)
    for t in (srctype, dsttype):
        new._node_frames[new.get_ntype_id(t)] = g._node_frames[g.get_ntype_id(t)].clone()
    return new


def sort_csr_by_tag(g, tag, tag_offset_name="_TAG_OFFSET", tag_type="node"):
    """Sort each row's out-neighbours by tag and store per-row tag offsets (reference sort_csr_by_tag)."""
    return _sort_by_tag(g, tag, tag_offset_name, tag_type, "csr")


def sort_csc_by_tag(g, tag, tag_offset_name="_TAG_OFFSET", tag_type="node"):
    return _sort_by_tag(g, tag, tag_offset_name, tag_type, "csc")


def _sort_by_tag(g, tag, name, tag_type, fmt):
    if len(g.etypes) > 1:
        raise DGLError("Only support homograph and bipartite graph")
    ug = g._graphs[0]
    indptr, indices, eids = ug.csr() if fmt == "csr" else ug.csc()
    nrow = indptr.shape[0] - 1
    e = eids if eids is not None else torch.arange(indices.shape[0], dtype=g.idtype, device=indices.device)
    tag = tag.to(indices.device).long()
    num_tags = int(tag.max()) + 1 if tag.numel() else 1
    tags = tag[indices.long()] if tag_type == "node" else tag[e.long()]
    rows = torch.repeat_interleave(torch.arange(nrow, device=indices.device), (indptr[1:] - indptr[:-1]).long())
    key = rows * num_tags + tags
    perm = torch.sort(key, stable=True)[1]
    new_indices, new_e = indices[perm], e[perm]
    cnt = torch.bincount(key, minlength=nrow * num_tags).view(nrow, num_tags)
    off = torch.zeros((nrow, num_tags + 1), dtype=g.idtype, device=indices.device)
    off[:, 1:] = torch.cumsum(cnt, 1)
    new = g.clone()
    if fmt == "csr":
        nug = UnitGraph.from_csr(ug.num_src, ug.num_dst, indptr, new_indices, new_e, formats=("csr",))
        new._graphs = [nug]
        new.srcdata[name] = off
    else:
        nug = UnitGraph.from_csc(ug.num_src, ug.num_dst, indptr, new_indices, new_e, formats=("csc",))
        new._graphs = [nug]
        new.dstdata[name] = off
    return new


def metis_perm(g, k):
    """Node permutation grouping the ``k`` partitions of :func:`metis_partition_assignment` (reference
    functional.py:3347); returns a numpy array like the reference."""
    from ..partition import metis_partition_assignment
    part = metis_partition_assignment(g if g.device.type == "cpu" else g.to("cpu"), k)
    return torch.sort(part, stable=True)[1].numpy().copy()


def rcmk_perm(g):
    """Reverse Cuthill-McKee node permutation (bandwidth reduction) of the symmetrised graph, computed natively
    (csrc/cpu/traversal.cc rcmk: pseudo-peripheral start per component, degree-ordered BFS) — reference
    functional.py:3371 calls scipy's ``reverse_cuthill_mckee``."""
    from .._native import loader
    from ..base import id_code
    cg = g if g.device.type == "cpu" else g.to("cpu")
    u, v = [t.contiguous() for t in cg.edges()]
    n = cg.num_nodes()
    perm = torch.empty(n, dtype=torch.int64)
    loader.cpu_lib().fn("dglb200_cpu_rcmk_coo", "ippllp")(id_code(cg.idtype), u.data_ptr(), v.data_ptr(), u.shape[0], n,
                                                          perm.data_ptr())
    return perm.numpy().copy()


def pairwise_squared_distance(x):
    """x: (n_samples, n_points, dims) -> (n_samples, n_points, n_points) squared Euclidean distances."""
    x2 = (x * x).sum(-1, keepdim=True)
    return x2 + x2.transpose(-1, -2) - 2 * x @ x.transpose(-1, -2)


NN_DESCENT_MIN_POINTS = 4096   # below this the exact kernels are cheaper than the descent rounds


def knn(k, x, x_segs, y=None, y_segs=None, algorithm="bruteforce", dist="euclidean"):
    """For every point of every segment of ``y`` the ``k`` nearest points of the same segment of ``x`` (self query when
    ``y`` is None). Returns a ``[2, k * len(y)]`` id tensor: row 0 indexes ``y``, row 1 indexes ``x`` (reference
    functional.py:641). 'kd-tree' (host) uses the native tree; the brute-force names use the shared-memory kernel on
    the GPU and the tree / GEMM path on the host; 'nn-descent' (self query) runs the approximate tensorised NN-descent for
    sets larger than ``NN_DESCENT_MIN_POINTS`` and the exact search below it."""
    if algorithm == "nn-descent" and y is not None:
        raise DGLError("Currently 'nn-descent' only supports self-query cases.")
    if y is None:
        y, y_segs = x, x_segs
    if x.device != y.device:
        raise DGLError("x and y must be on the same device")
    x_segs = torch.as_tensor(x_segs).long().cpu()
    y_segs = torch.as_tensor(y_segs).long().cpu()
    if x_segs.numel() != y_segs.numel():
        raise DGLError("x_segs and y_segs must describe the same number of point sets")
    lo = int(x_segs.min()) if x_segs.numel() else 0
    if k > lo:
        dgl_warning(f"'k' should be less than or equal to the number of points in 'x'expect k <= {lo}, got k = {k}, "
                    f"use k = {lo}")
        k = lo
    if k <= 0:
        raise DGLError(f"Invalid k value. expect k > 0, got k = {k}")
    if x.shape[0] == 0 or y.shape[0] == 0:
        raise DGLError("Find empty point set")
    dist = dist.lower()
    if dist not in ("euclidean", "cosine"):
        raise DGLError(f"Only ['euclidean', 'cosine'] are supported for distancecomputation, got {dist}")
    if dist == "cosine":
        x = x / (x.norm(dim=1, keepdim=True) + 1e-5)
        y = y / (y.norm(dim=1, keepdim=True) + 1e-5)
    xo = torch.cat([torch.zeros(1, dtype=torch.long), x_segs.cumsum(0)])
    yo = torch.cat([torch.zeros(1, dtype=torch.long), y_segs.cumsum(0)])
    if algorithm == "nn-descent" and x.shape[0] > NN_DESCENT_MIN_POINTS:
        # approximate search for large sets (any dimension, any device): tensorised NN-descent, transforms/nn_descent.py
        from .nn_descent import nn_descent
        idx = nn_descent(x, xo, k)
    elif x.device.type == "cpu" and x.dtype == y.dtype and x.dtype in (torch.float32, torch.float64) and \
            (algorithm == "kd-tree" or x.shape[1] <= 16):
        idx = kdtree_knn(x, y, k, xo, yo)[0]
    elif x.is_cuda and x.shape[1] <= 16 and k <= 64 and x.dtype == y.dtype and x.dtype in (
            torch.float32, torch.float16, torch.bfloat16):
        from .._graph_kernels import knn_brute
        idx = knn_brute(x, xo.to(x.device), y, yo.to(x.device), k)
    else:
        idx = torch.cat([_pairwise_topk(x[xo[s]:xo[s + 1]], y[yo[s]:yo[s + 1]], k) + int(xo[s])
                         for s in range(x_segs.numel())])
    rows = torch.arange(y.shape[0], device=idx.device).view(-1, 1).expand_as(idx)
    return torch.stack([rows.reshape(-1), idx.reshape(-1)]).to(torch.int64)


def reorder_graph(g, node_permute_algo=None, edge_permute_algo="src", store_ids=True, permute_config=None):
    """Reorder nodes ('rcmk', 'metis', 'custom') and edges ('src', 'dst', 'custom')
    (reference reorder_graph)."""
    if not g.is_homogeneous:
        raise DGLError("Only homogeneous graphs are supported.")
    permute_config = permute_config or {}
    n = g.num_nodes()
    if node_permute_algo is None:
        perm = torch.arange(n, device=g.device)
    elif node_permute_algo == "rcmk":
        perm = torch.as_tensor(rcmk_perm(g)).to(g.device)
    elif node_permute_algo == "metis":
        k = permute_config.get("k")
        if k is None:
            raise DGLError("Partition parts 'k' is required for metis.")
        perm = torch.as_tensor(metis_perm(g, k)).to(g.device)
    elif node_permute_algo == "custom":
        perm = _to_tensor(permute_config["nodes_perm"], torch.int64, g.device)
        if perm.shape[0] != n:
            raise DGLError("Expect nodes_perm to have the same size as the number of nodes")
    else:
        raise DGLError(f"Unexpected node_permute_algo: {node_permute_algo}")
    perm = perm.long()
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device=perm.device)
    u, v = g.all_edges()
    nu, nv = inv[u.long()].to(g.idtype), inv[v.long()].to(g.idtype)
    if edge_permute_algo == "src":
        eperm = torch.sort(nu.long() * n + nv.long(), stable=True)[1]
    elif edge_permute_algo == "dst":
        eperm = torch.sort(nv.long() * n + nu.long(), stable=True)[1]
    elif edge_permute_algo == "custom":
        eperm = _to_tensor(permute_config["edges_perm"], torch.int64, g.device)
    else:
        raise DGLError(f"Unexpected edge_permute_algo: {edge_permute_algo}")
    new = make_graph((nu[eperm], nv[eperm]), num_nodes=n, idtype=g.idtype, device=g.device)
    for k, val in g.ndata.items():
        new.ndata[k] = val[perm]
    for k, val in g.edata.items():
        new.edata[k] = val[eperm]
    if store_ids:
        new.ndata[NID] = perm.to(g.idtype)
        new.edata[EID] = eperm.to(g.idtype)
    return new


def to_bidirected_coo(g, e_weights=None):
    """(u, v, w) of the symmetrised edge set without self loops (weights duplicated for both directions)."""
    u, v = g.edges()
    m = u != v
    u, v = u[m], v[m]
    w = None
    if e_weights is not None:
        w = e_weights.to(u.device)[m]
        w = torch.cat([w, w])
    return torch.cat([u, v]).contiguous(), torch.cat([v, u]).contiguous(), w


def kdtree_knn(x, y, k, x_offsets=None, y_offsets=None, return_distance=False):
    """k nearest rows of ``x`` for every row of ``y`` with a native kd-tree (host tensors, float32/64). With offsets the
    search is segmented: queries of segment s only see data points of segment s. Returns (idx [len(y), k], dist|None);
    rows with fewer than k candidates are padded with -1 (reference src/graph/transform/cpu/knn.cc KdTreeKNN)."""
    from .._native import loader
    from ..base import dtype_code
    x, y = x.contiguous(), y.contiguous()
    xo = torch.tensor([0, x.shape[0]], dtype=torch.int64) if x_offsets is None else torch.as_tensor(x_offsets).long().contiguous()
    yo = torch.tensor([0, y.shape[0]], dtype=torch.int64) if y_offsets is None else torch.as_tensor(y_offsets).long().contiguous()
    idx = torch.empty((y.shape[0], k), dtype=torch.int64)
    dist = torch.empty((y.shape[0], k), dtype=x.dtype) if return_distance else None
    loader.cpu_lib().fn("dglb200_cpu_knn_kdtree", "ipppplllpp")(
        dtype_code(x.dtype), x.data_ptr(), xo.data_ptr(), y.data_ptr(), yo.data_ptr(), xo.shape[0] - 1, x.shape[1], int(k),
        idx.data_ptr(), 0 if dist is None else dist.data_ptr())
    return idx, dist


def _pairwise_topk(x, y, k, exclude_self_offset=None, dist="euclidean", chunk=4096):
    """for each row of y find k nearest rows of x. returns idx [len(y), k]"""
    if x.is_cuda and x.shape[1] <= 16 and k <= 64 and dist == "euclidean" and x.dtype == y.dtype \
            and x.dtype in (torch.float32, torch.float16, torch.bfloat16):
        # low-dimensional point clouds: shared-memory brute force, one thread per query (csrc/cuda/graph_kernels.cu)
        from .._graph_kernels import knn_brute
        doff = torch.tensor([0, x.shape[0]], dtype=torch.int64, device=x.device)
        qoff = torch.tensor([0, y.shape[0]], dtype=torch.int64, device=x.device)
        return knn_brute(x, doff, y, qoff, k)
    if dist == "cosine":
        x = torch.nn.functional.normalize(x, dim=-1)
        y = torch.nn.functional.normalize(y, dim=-1)
    if x.device.type == "cpu" and x.shape[1] <= 16 and x.dtype == y.dtype and x.dtype in (torch.float32, torch.float64) \
            and x.shape[0] > 256:
        # host point clouds: kd-tree (csrc/cpu/spatial.cc); on unit vectors Euclidean order == cosine order
        return kdtree_knn(x, y, k)[0]
    outs = []
    for s in range(0, y.shape[0], chunk):
        yc = y[s:s + chunk]
        if dist == "cosine":
            d = 1 - yc @ x.t()
        else:
            d = torch.cdist(yc, x)
        outs.append(torch.topk(d, k, dim=1, largest=False)[1])
    return torch.cat(outs, 0) if outs else torch.empty((0, k), dtype=torch.int64, device=x.device)


def knn_graph(x, k, algorithm="bruteforce-blas", dist="euclidean", exclude_self=False):
    """k-nearest-neighbour graph: edge j -> i iff j is among i's k nearest points (reference knn_graph).
    All algorithm names map to a chunked GEMM-based brute force here (tensor-core friendly on B200)."""
    if x.dim() == 2:
        x = x.unsqueeze(0)
    B, n, _ = x.shape
    kk = k + 1 if exclude_self else k
    srcs, dsts = [], []
    for b in range(B):
        idx = _pairwise_topk(x[b], x[b], min(kk, n), dist=dist)
        dst = torch.arange(n, device=x.device).view(-1, 1).expand_as(idx)
        s, d = idx.reshape(-1), dst.reshape(-1)
        if exclude_self:
            m = s != d
            # keep exactly k per row: drop self if present else drop farthest
            keep = torch.ones_like(idx, dtype=torch.bool)
            is_self = idx == torch.arange(n, device=x.device).view(-1, 1)
            has_self = is_self.any(1)
            keep[is_self] = False
            keep[~has_self, -1] = False
            s, d = idx[keep], dst[keep]
        srcs.append(s + b * n)
        dsts.append(d + b * n)
    g = make_graph((torch.cat(srcs), torch.cat(dsts)), num_nodes=B * n, device=x.device)
    if B > 1:
        g.set_batch_num_nodes(torch.full((B,), n, dtype=g.idtype, device=x.device))
        g.set_batch_num_edges(torch.full((B,), g.num_edges() // B, dtype=g.idtype, device=x.device))
    return g


def segmented_knn_graph(x, k, segs, algorithm="bruteforce-blas", dist="euclidean", exclude_self=False):
    segs = [int(n) for n in segs]
    kk = k + 1 if exclude_self else k
    if x.device.type == "cpu" and x.shape[1] <= 16 and x.dtype in (torch.float32, torch.float64) and segs \
            and min(segs) >= kk:
        # one native call: a kd-tree per segment, queries confined to their own segment
        pts = torch.nn.functional.normalize(x, dim=-1) if dist == "cosine" else x
        offs = torch.tensor([0] + segs).cumsum(0)
        idx, _ = kdtree_knn(pts, pts, kk, offs, offs)
        n = x.shape[0]
        dst = torch.arange(n).view(-1, 1).expand_as(idx)
        if exclude_self:
            keep = torch.ones_like(idx, dtype=torch.bool)
            is_self = idx == dst
            keep[is_self] = False
            keep[~is_self.any(1), -1] = False
            return make_graph((idx[keep], dst[keep]), num_nodes=n)
        return make_graph((idx.reshape(-1), dst.reshape(-1)), num_nodes=n)
    off = 0
    srcs, dsts = [], []
    for n in segs:
        sub = x[off: off + n]
        g = knn_graph(sub, min(k, n), dist=dist, exclude_self=exclude_self)
        u, v = g.all_edges()
        srcs.append(u + off)
        dsts.append(v + off)
        off += n
    g = make_graph((torch.cat(srcs), torch.cat(dsts)), num_nodes=off, device=x.device)
    return g


def radius_graph(x, r, p=2, self_loop=False, compute_mode="donot_use_mm_for_euclid_dist", get_distances=False):
    d = torch.cdist(x, x, p=p)
    m = d <= r
    if not self_loop:
        m.fill_diagonal_(False)
    # edge j -> i for j within radius of i
    dst, src = torch.nonzero(m, as_tuple=True)
    g = make_graph((src, dst), num_nodes=x.shape[0], device=x.device)
    if get_distances:
        return g, d[dst, src].unsqueeze(-1)
    return g


def random_walk_pe(g, k, eweight_name=None):
    """Random-walk positional encoding: diag of (D^-1 A)^i for i=1..k (reference random_walk_pe)."""
    import scipy.sparse as sp
    n = g.num_nodes()
    u, v = g.all_edges()
    w = g.edata[eweight_name].cpu().numpy().astype(float) if eweight_name else np.ones(u.shape[0])
    a = sp.coo_matrix((w, (u.cpu().numpy(), v.cpu().numpy())), shape=(n, n)).tocsr()
    deg = np.asarray(a.sum(1)).reshape(-1)
    deg[deg == 0] = 1
    rw = sp.diags(1.0 / deg) @ a
    pe = [rw.diagonal()]
    cur = rw
    for _ in range(k - 1):
        cur = cur @ rw
        pe.append(cur.diagonal())
    return torch.from_numpy(np.stack(pe, -1)).float().to(g.device)


def lap_pe(g, k, padding=False, return_eigval=False):
    """Laplacian positional encoding: k smallest non-trivial eigenvectors of the sym-normalised Laplacian."""
    import scipy.sparse as sp
    n = g.num_nodes()
    if not padding and n <= k:
        raise DGLError(f"the number of eigenvectors k must be smaller than the number of nodes n, {k} and {n} detected.")
    a = _scipy_adj(g).astype(float)
    deg = np.asarray(g.in_degrees().cpu().numpy(), dtype=float).clip(1)
    nrm = sp.diags(deg ** -0.5)
    lap = sp.eye(n) - nrm @ a @ nrm
    evals, evecs = np.linalg.eigh(lap.toarray())
    kk = min(n - 1, k)
    order = np.argsort(evals)[1: kk + 1]
    topv, tope = evecs[:, order].real, evals[order].real
    rand_sign = 2 * (np.random.rand(kk) > 0.5) - 1.0
    pe = torch.from_numpy(rand_sign * topv).float()
    ev = torch.from_numpy(tope).float()
    if n <= k:
        pe = torch.cat([pe, torch.zeros(n, k - n + 1)], 1)
        ev = torch.cat([ev, torch.full((k - n + 1,), float("nan"))])
    pe = pe.to(g.device)
    if return_eigval:
        return pe, ev.to(g.device)
    return pe


laplacian_pe = lap_pe


def double_radius_node_labeling(g, src, dst):
    """DRNL labels for link (src, dst) as in SEAL (reference double_radius_node_labeling)."""
    from scipy.sparse.csgraph import shortest_path
    adj = _scipy_adj(g)
    adj = ((adj + adj.T) > 0).astype(float).tocsr()
    src, dst = (dst, src) if src > dst else (src, dst)
    idx = list(range(src)) + list(range(src + 1, adj.shape[0]))
    adj_wo_src = adj[idx, :][:, idx]
    idx = list(range(dst)) + list(range(dst + 1, adj.shape[0]))
    adj_wo_dst = adj[idx, :][:, idx]
    d2s = shortest_path(adj_wo_dst, directed=False, unweighted=True, indices=src)
    d2s = np.insert(d2s, dst, 0, axis=0)
    d2s = torch.from_numpy(d2s)
    d2d = shortest_path(adj_wo_src, directed=False, unweighted=True, indices=dst - 1)
    d2d = np.insert(d2d, src, 0, axis=0)
    d2d = torch.from_numpy(d2d)
    dist = d2s + d2d
    dist_over_2, dist_mod_2 = torch.div(dist, 2, rounding_mode="floor"), dist % 2
    z = 1 + torch.min(d2s, d2d)
    z += dist_over_2 * (dist_over_2 + dist_mod_2 - 1)
    z[src] = 1.0
    z[dst] = 1.0
    z[torch.isnan(z)] = 0.0
    z[torch.isinf(z)] = 0.0
    return z.to(torch.long)


def shortest_dist(g, root=None, return_paths=False):
    """BFS shortest distances (all pairs when root is None) (reference shortest_dist)."""
    from scipy.sparse.csgraph import shortest_path
    adj = _scipy_adj(g)
    if root is None:
        d, pred = shortest_path(adj, directed=True, unweighted=True, return_predecessors=True)
    else:
        d, pred = shortest_path(adj, directed=True, unweighted=True, return_predecessors=True, indices=root)
    d = np.where(np.isinf(d), -1, d)
    dist = torch.from_numpy(d).long().to(g.device)
    if not return_paths:
        return dist
    # reconstruct edge-id paths
    n = g.num_nodes()
    maxlen = int(dist.max()) if dist.numel() else 0
    u, v = g.all_edges()
    emap = {}
    for e, (a, b) in enumerate(zip(u.tolist(), v.tolist())):
        emap.setdefault((a, b), e)

    def path(src, pr):
        out = torch.full((n, maxlen), -1, dtype=torch.long)
        for t in range(n):
            if pr[t] < 0 or t == src:
                continue
            nodes = [t]
            while nodes[-1] != src and pr[nodes[-1]] >= 0:
                nodes.append(int(pr[nodes[-1]]))
            nodes = nodes[::-1]
            for i in range(len(nodes) - 1):
                out[t, i] = emap[(nodes[i], nodes[i + 1])]
        return out

    if root is None:
        paths = torch.stack([path(s, pred[s]) for s in range(n)])
    else:
        paths = path(root, pred)
    return dist, paths.to(g.device)


def svd_pe(g, k, padding=False, random_flip=True):
    n = g.num_nodes()
    if not padding and n < k:
        raise DGLError(f"The number of singular values k must be no greater than the number of nodes n, {k} and {n} detected.")
    a = torch.from_numpy(_scipy_adj(g).toarray()).float()
    u, d, vh = torch.linalg.svd(a)
    kk = min(n, k)
    m = min(kk, d.shape[0])
    topd = d[:m]
    topu = u[:, :m]
    topv = vh.t()[:, :m]
    ds = torch.sqrt(topd)
    if random_flip:
        sign = 2 * (torch.rand(m) > 0.5).float() - 1
    else:
        sign = torch.ones(m)
    enc_u = sign * topu * ds
    enc_v = sign * topv * ds
    if n < k:
        pad = torch.zeros(n, k - n)
        enc_u = torch.cat([enc_u, pad], 1)
        enc_v = torch.cat([enc_v, pad], 1)
    return torch.cat([enc_u, enc_v], 1).to(g.device)


def _cast_feats(g, dtype):
    new = g.clone()
    for fr in new._node_frames + new._edge_frames:
        for k in list(fr.keys()):
            v = fr[k]
            if torch.is_tensor(v) and v.is_floating_point():
                fr[k] = v.to(dtype)
    return new


def to_half(g):
    return _cast_feats(g, torch.float16)


def to_bfloat16(g):
    """Shallow copy whose floating-point node / edge features are bfloat16 (reference functional.py:3729)."""
    return _cast_feats(g, torch.bfloat16)


def to_float(g):
    return _cast_feats(g, torch.float32)


def to_double(g):
    return _cast_feats(g, torch.float64)


def adj_product_graph(A, B, weight_name, etype="_E"):
    """Graph of A @ B with weights (reference adj_product_graph); uses dgl_b200.sparse.spspmm."""
    from ..sparse import spmatrix, spspmm
    ua, va = A.all_edges()
    ub, vb = B.all_edges()
    sa, _, da = A.canonical_etypes[0]
    sb, _, db = B.canonical_etypes[0]
    ma = spmatrix(torch.stack([ua, va]), A.edata[weight_name], (A.num_nodes(sa), A.num_nodes(da)))
    mb = spmatrix(torch.stack([ub, vb]), B.edata[weight_name], (B.num_nodes(sb), B.num_nodes(db)))
    mc = spspmm(ma, mb)
    r, c = mc.coo()
    if sa == db:
        g = make_graph((r, c), num_nodes=max(A.num_nodes(sa), B.num_nodes(db)), idtype=A.idtype, device=A.device)
    else:
        g = make_heterograph({(sa, etype, db): (r, c)}, {sa: A.num_nodes(sa), db: B.num_nodes(db)}, idtype=A.idtype,
                             device=A.device)
    g.edata[weight_name] = mc.val
    return g


def adj_sum_graph(graphs, weight_name):
    from ..sparse import spmatrix
    g0 = graphs[0]
    s, _, d = g0.canonical_etypes[0]
    acc = None
    for g in graphs:
        u, v = g.all_edges()
        m = spmatrix(torch.stack([u, v]), g.edata[weight_name], (g.num_nodes(s), g.num_nodes(d)))
        acc = m if acc is None else acc + m
    acc = acc.coalesce()
    r, c = acc.coo()
    if s == d:
        out = make_graph((r, c), num_nodes=g0.num_nodes(s), idtype=g0.idtype, device=g0.device)
    else:
        out = make_heterograph({g0.canonical_etypes[0]: (r, c)}, {s: g0.num_nodes(s), d: g0.num_nodes(d)},
                               idtype=g0.idtype, device=g0.device)
    out.edata[weight_name] = acc.val
    return out


def out_subgraph_like(g, nodes):
    return g.out_subgraph(nodes)


def norm_by_dst(g, etype=None):
    """1/in-degree per edge (reference functional.py:norm_by_dst)."""
    _, v = g.all_edges(etype=etype)
    deg = g.in_degrees(etype=etype).float().clamp(min=1)
    return 1.0 / deg[v.long()]


def metis_partition_assignment(g, k, balance_ntypes=None, balance_edges=False, mode="k-way", objtype="cut"):
    from ..partition import metis_partition_assignment as f
    return f(g, k, balance_ntypes=balance_ntypes, balance_edges=balance_edges, mode=mode, objtype=objtype)


def metis_partition(g, k, extra_cached_hops=0, reshuffle=False, balance_ntypes=None, balance_edges=False, mode="k-way"):
    from ..partition import metis_partition as f
    return f(g, k, extra_cached_hops=extra_cached_hops, reshuffle=reshuffle, balance_ntypes=balance_ntypes,
             balance_edges=balance_edges, mode=mode)


def partition_graph_with_halo(g, node_part, extra_cached_hops, reshuffle=False):
    from ..partition import partition_graph_with_halo as f
    return f(g, node_part, extra_cached_hops, reshuffle=reshuffle)
