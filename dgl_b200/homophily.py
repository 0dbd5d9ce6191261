"""Homophily measures (parity: /root/reference/python/dgl/homophily.py)."""
from __future__ import annotations

import torch

from . import function as fn

__all__ = ["node_homophily", "edge_homophily", "linkx_homophily", "adjusted_homophily"]


def check_pytorch():
    """The measures are written against torch tensors (the reference guards on its backend switch here)."""
    return True


def get_long_edges(graph):
    """(src, dst) of every edge as int64 tensors."""
    src, dst = graph.edges()
    return src.long(), dst.long()


def node_homophily(graph, y):
    with graph.local_scope():
        src, dst = graph.edges()
        graph.edata["same_class"] = (y[src.long()] == y[dst.long()]).float()
        graph.update_all(fn.copy_e("same_class", "m"), fn.mean("m", "same_class_deg"))
        return graph.ndata["same_class_deg"].mean(dim=0).item()


def edge_homophily(graph, y):
    src, dst = graph.edges()
    return (y[src.long()] == y[dst.long()]).float().mean().item()


def linkx_homophily(graph, y):
    with graph.local_scope():
        src, dst = graph.edges()
        graph.edata["same_class"] = (y[src.long()] == y[dst.long()]).float()
        graph.update_all(fn.copy_e("same_class", "m"), fn.sum("m", "same_class_deg"))
        deg = graph.in_degrees().float()
        num_nodes = graph.num_nodes()
        num_classes = int(y.max()) + 1
        value = torch.tensor(0.0, device=graph.device)
        for k in range(num_classes):
            class_mask = y == k
            same_class_deg_k = graph.ndata["same_class_deg"][class_mask].sum()
            deg_k = deg[class_mask].sum()
            num_nodes_k = class_mask.sum()
            value = value + torch.clamp(same_class_deg_k / deg_k.clamp(min=1) - num_nodes_k / num_nodes, min=0)
        return (value / max(num_classes - 1, 1)).item()


def adjusted_homophily(graph, y):
    edge_homo = edge_homophily(graph, y)
    degrees = graph.in_degrees().float()
    num_classes = int(y.max()) + 1
    degree_sums = torch.zeros(num_classes, device=graph.device).index_add_(0, y.long(), degrees)
    adjust = (degree_sums ** 2).sum() / graph.num_edges() ** 2
    return ((edge_homo - adjust) / (1 - adjust)).item()
