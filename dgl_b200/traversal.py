"""Graph traversal frontier generators (parity: /root/reference/python/dgl/traversal.py, src/graph/traversal.cc):
bfs_nodes_generator, bfs_edges_generator, topological_nodes_generator, dfs_edges_generator, dfs_labeled_edges_generator."""
from __future__ import annotations

import torch

from ._native import loader
from .base import DGLError, id_code
from .heterograph import _to_tensor

__all__ = ["bfs_nodes_generator", "bfs_edges_generator", "topological_nodes_generator", "dfs_edges_generator",
           "dfs_labeled_edges_generator"]


def _csr(g, reverse):
    ug = g._graphs[0]
    return ug.csc() if reverse else ug.csr()


def _split(ids, sections):
    return list(torch.split(ids, sections.tolist())) if sections.numel() else []


def _native_bfs(graph, source, reverse, want_edges):
    """Host graphs: one native pass (csrc/cpu/traversal.cc); frontiers are in discovery order like the reference."""
    indptr, indices, eids = _csr(graph, reverse)
    n = graph.num_nodes()
    src = _to_tensor(source, graph.idtype, graph.device).contiguous()
    out = torch.empty(n, dtype=graph.idtype)
    sections = torch.empty(n + 1, dtype=torch.int64)
    counts = torch.zeros(2, dtype=torch.int64)
    loader.cpu_lib().fn("dglb200_cpu_bfs", "ippplplipppp")(
        id_code(graph.idtype), indptr.data_ptr(), indices.data_ptr(), 0 if eids is None else eids.data_ptr(), n,
        src.data_ptr(), src.shape[0], int(want_edges), out.data_ptr(), sections.data_ptr(), counts.data_ptr(),
        counts[1:].data_ptr())
    return _split(out[:int(counts[1])], sections[:int(counts[0])])


def bfs_nodes_generator(graph, source, reverse=False):
    """Frontiers of a BFS from ``source`` (native on the host, vectorised level-synchronous expansion on the GPU)."""
    if graph.device.type == "cpu":
        return _native_bfs(graph, source, reverse, False)
    indptr, indices, _ = _csr(graph, reverse)
    n = graph.num_nodes()
    frontier = _to_tensor(source, graph.idtype, graph.device)
    seen = torch.zeros(n, dtype=torch.bool, device=graph.device)
    seen[frontier.long()] = True
    out = [frontier]
    while True:
        s = indptr[frontier.long()].long()
        cnt = indptr[frontier.long() + 1].long() - s
        pos = torch.repeat_interleave(s, cnt) + (torch.arange(int(cnt.sum()), device=s.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
        nb = torch.unique(indices[pos])
        nb = nb[~seen[nb.long()]]
        if nb.numel() == 0:
            break
        seen[nb.long()] = True
        out.append(nb)
        frontier = nb
    return out


def bfs_edges_generator(graph, source, reverse=False):
    """Edge frontiers of a BFS tree (one tree edge per newly discovered node)."""
    if graph.device.type == "cpu":
        return _native_bfs(graph, source, reverse, True)
    indptr, indices, eids = _csr(graph, reverse)
    n = graph.num_nodes()
    frontier = _to_tensor(source, graph.idtype, graph.device)
    seen = torch.zeros(n, dtype=torch.bool, device=graph.device)
    seen[frontier.long()] = True
    out = []
    while True:
        s = indptr[frontier.long()].long()
        cnt = indptr[frontier.long() + 1].long() - s
        pos = torch.repeat_interleave(s, cnt) + (torch.arange(int(cnt.sum()), device=s.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
        nb = indices[pos]
        e = eids[pos] if eids is not None else pos.to(graph.idtype)
        new = ~seen[nb.long()]
        nb, e = nb[new], e[new]
        if nb.numel() == 0:
            break
        uniq, inv = torch.unique(nb, return_inverse=True)
        first = torch.full((uniq.shape[0],), nb.shape[0], dtype=torch.long, device=nb.device)
        first.scatter_reduce_(0, inv, torch.arange(nb.shape[0], device=nb.device), reduce="amin")
        out.append(e[first])
        seen[uniq.long()] = True
        frontier = uniq
    return out


def topological_nodes_generator(graph, reverse=False):
    """Kahn levels of a DAG."""
    indptr, indices, _ = _csr(graph, reverse)
    n = graph.num_nodes()
    if graph.device.type == "cpu":
        out = torch.empty(n, dtype=graph.idtype)
        sections = torch.empty(n + 1, dtype=torch.int64)
        counts = torch.zeros(2, dtype=torch.int64)
        loader.cpu_lib().fn("dglb200_cpu_topological", "ipplpppp")(
            id_code(graph.idtype), indptr.data_ptr(), indices.data_ptr(), n, out.data_ptr(), sections.data_ptr(),
            counts.data_ptr(), counts[1:].data_ptr())
        if int(counts[1]) != n:
            raise DGLError("The graph has cycles; topological traversal requires a DAG.")
        return _split(out, sections[:int(counts[0])])
    deg = (graph.out_degrees() if reverse else graph.in_degrees()).clone().long()
    frontier = torch.nonzero(deg == 0).squeeze(1).to(graph.idtype)
    out, visited = [], 0
    while frontier.numel():
        out.append(frontier)
        visited += frontier.numel()
        s = indptr[frontier.long()].long()
        cnt = indptr[frontier.long() + 1].long() - s
        pos = torch.repeat_interleave(s, cnt) + (torch.arange(int(cnt.sum()), device=s.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
        nb = indices[pos].long()
        deg.index_add_(0, nb, -torch.ones_like(nb))
        cand = torch.unique(nb)
        frontier = cand[deg[cand] == 0].to(graph.idtype)
    if visited != n:
        raise DGLError("The graph has cycles; topological traversal requires a DAG.")
    return out


def _dfs(graph, source, reverse, has_reverse_edge, has_nontree_edge):
    """Per-source labelled edge streams (FORWARD=0, REVERSE=1, NONTREE=2), native (csrc/cpu/traversal.cc)."""
    indptr, indices, eids = [t.cpu().contiguous() if t is not None else None for t in _csr(graph, reverse)]
    n = graph.num_nodes()
    cap = int(indices.shape[0]) + 2 * n + 1
    fn = loader.cpu_lib().fn("dglb200_cpu_dfs", "ippplliipplp")
    all_edges, all_labels = [], []
    for src in _to_tensor(source, graph.idtype, torch.device("cpu")).tolist():
        edges = torch.empty(cap, dtype=indptr.dtype)
        labels = torch.empty(cap, dtype=torch.int64)
        count = torch.zeros(1, dtype=torch.int64)
        fn(id_code(indptr.dtype), indptr.data_ptr(), indices.data_ptr(), 0 if eids is None else eids.data_ptr(), n,
           int(src), int(has_reverse_edge), int(has_nontree_edge), edges.data_ptr(), labels.data_ptr(), cap,
           count.data_ptr())
        k = int(count)
        all_edges.append(edges[:k].to(graph.idtype))
        all_labels.append(labels[:k])
    return all_edges, all_labels


def _merge_traversals(streams):
    """Frontier i holds the i-th element of every stream that is long enough (reference
    src/graph/traversal.cc MergeMultipleTraversals / ComputeMergedSections)."""
    if not streams:
        return []
    if len(streams) == 1:
        return list(streams[0].split(1))
    longest = max(int(t.shape[0]) for t in streams)
    return [torch.stack([t[i] for t in streams if t.shape[0] > i]) for i in range(longest)]


def dfs_edges_generator(graph, source, reverse=False):
    edges, _ = _dfs(graph, source, reverse, False, False)
    return _merge_traversals(edges)


def dfs_labeled_edges_generator(graph, source, reverse=False, has_reverse_edge=False, has_nontree_edge=False,
                                return_labels=True):
    edges, labels = _dfs(graph, source, reverse, has_reverse_edge, has_nontree_edge)
    fe = _merge_traversals(edges)
    if return_labels:
        return fe, _merge_traversals(labels)
    return fe
