"""Graph traversal frontier generators (parity: /root/reference/python/dgl/traversal.py, src/graph/traversal.cc):
bfs_nodes_generator, bfs_edges_generator, topological_nodes_generator, dfs_edges_generator, dfs_labeled_edges_generator."""
from __future__ import annotations

import torch

from .base import DGLError
from .heterograph import _to_tensor

__all__ = ["bfs_nodes_generator", "bfs_edges_generator", "topological_nodes_generator", "dfs_edges_generator",
           "dfs_labeled_edges_generator"]


def _csr(g, reverse):
    ug = g._graphs[0]
    return ug.csc() if reverse else ug.csr()


def bfs_nodes_generator(graph, source, reverse=False):
    """Frontiers of a BFS from ``source`` (vectorised level-synchronous expansion)."""
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
    FORWARD, REVERSE, NONTREE = 0, 1, 2
    indptr, indices, eids = [t.cpu() if t is not None else None for t in _csr(graph, reverse)]
    ip, ix = indptr.tolist(), indices.tolist()
    ei = eids.tolist() if eids is not None else list(range(len(ix)))
    src_list = _to_tensor(source, graph.idtype, torch.device("cpu")).tolist()
    all_edges, all_labels = [], []
    for src in src_list:
        visited = {src}
        stack = [(src, ip[src])]
        edges, labels = [], []
        while stack:
            u, p = stack[-1]
            if p < ip[u + 1]:
                stack[-1] = (u, p + 1)
                v = ix[p]
                if v not in visited:
                    visited.add(v)
                    edges.append(ei[p]); labels.append(FORWARD)
                    stack.append((v, ip[v]))
                elif has_nontree_edge:
                    edges.append(ei[p]); labels.append(NONTREE)
            else:
                stack.pop()
                if has_reverse_edge and stack:
                    pu, pp = stack[-1]
                    edges.append(ei[pp - 1]); labels.append(REVERSE)
        all_edges.append(edges)
        all_labels.append(labels)
    return all_edges, all_labels


def dfs_edges_generator(graph, source, reverse=False):
    edges, _ = _dfs(graph, source, reverse, False, False)
    flat = [e for es in edges for e in es]
    return [torch.tensor([e], dtype=graph.idtype) for e in flat]


def dfs_labeled_edges_generator(graph, source, reverse=False, has_reverse_edge=False, has_nontree_edge=False,
                                return_labels=True):
    edges, labels = _dfs(graph, source, reverse, has_reverse_edge, has_nontree_edge)
    fe = [torch.tensor([e], dtype=graph.idtype) for es in edges for e in es]
    fl = [torch.tensor([l], dtype=torch.int64) for ls in labels for l in ls]
    if return_labels:
        return fe, fl
    return fe
