"""Homophily measures of a node-labelled graph.

Definitions follow /root/reference/python/dgl/homophily.py:27-270 (node / edge / LINKX class-insensitive / adjusted
homophily). All four are computed from ONE pair of per-edge tensors (endpoint labels) with scatter reductions over class
and node ids - no message-passing round trip and no Python loop over classes."""
from __future__ import annotations

import torch

__all__ = ["node_homophily", "edge_homophily", "linkx_homophily", "adjusted_homophily"]


def check_pytorch():
    """Kept for API parity (the reference guards on its backend switch); this package is torch-only."""
    return True


def get_long_edges(graph):
    u, v = graph.edges()
    return u.long(), v.long()


def _agreement(graph, y):
    """(dst ids, 1.0 where an edge joins two nodes of the same class else 0.0)."""
    u, v = get_long_edges(graph)
    return v, (y[u] == y[v]).to(torch.float32)


def edge_homophily(graph, y):
    """Fraction of edges whose endpoints share a label."""
    return _agreement(graph, y)[1].mean().item()


def node_homophily(graph, y):
    """Mean over nodes of the fraction of in-neighbours with the node's own label (nodes without in-edges count 0)."""
    v, same = _agreement(graph, y)
    n = graph.num_nodes()
    hits = torch.zeros(n, device=same.device).index_add_(0, v, same)
    deg = torch.zeros(n, device=same.device).index_add_(0, v, torch.ones_like(same))
    return (hits / deg.clamp(min=1)).mean().item()


def linkx_homophily(graph, y):
    """Class-insensitive homophily of Lim et al.: (1 / (C - 1)) * sum_k max(0, h_k - |C_k| / N), h_k the fraction of the
    in-edges of class-k nodes that come from class k."""
    v, same = _agreement(graph, y)
    y = y.long()
    num_classes = int(y.max()) + 1
    cls_of_dst = y[v]
    same_per_class = torch.zeros(num_classes, device=same.device).index_add_(0, cls_of_dst, same)
    edges_per_class = torch.zeros(num_classes, device=same.device).index_add_(0, cls_of_dst, torch.ones_like(same))
    share = torch.bincount(y, minlength=num_classes).to(torch.float32) / graph.num_nodes()
    excess = (same_per_class / edges_per_class.clamp(min=1) - share.to(same.device)).clamp(min=0)
    return (excess.sum() / max(num_classes - 1, 1)).item()


def adjusted_homophily(graph, y):
    """Edge homophily corrected for the class degree distribution: (h_edge - sum_k D_k^2 / (E)^2) / (1 - sum_k D_k^2 / E^2)
    with D_k the total degree of class k (Platonov et al.). Defined on the undirected simple graph: the input is symmetrised
    and de-duplicated first."""
    from .transforms import to_bidirected
    graph = to_bidirected(graph.cpu()).to(y.device)
    v, same = _agreement(graph, y)
    num_classes = int(y.max()) + 1
    class_degree = torch.zeros(num_classes, device=same.device).index_add_(0, y.long()[v], torch.ones_like(same))
    expected = (class_degree ** 2).sum() / float(graph.num_edges()) ** 2
    return ((same.mean() - expected) / (1 - expected)).item()
