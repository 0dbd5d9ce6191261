"""Merge graphs sharing a node space (parity: /root/reference/python/dgl/merge.py)."""
from __future__ import annotations

import torch

from .base import DGLError
from .convert import heterograph

__all__ = ["merge"]


def merge(graphs):
    if len(graphs) == 0:
        raise DGLError("The list of graphs must not be empty.")
    ref = graphs[0]
    ntypes, cets = ref.ntypes, ref.canonical_etypes
    num_nodes = {nt: max(g.num_nodes(nt) for g in graphs) for nt in ntypes}
    data = {}
    for c in cets:
        us, vs = zip(*[g.all_edges(etype=c) for g in graphs])
        data[c] = (torch.cat(us), torch.cat(vs))
    out = heterograph(data, num_nodes, idtype=ref.idtype, device=ref.device)
    for nt in ntypes:
        for k in ref.nodes[nt].data.keys():
            best = max(graphs, key=lambda g: g.num_nodes(nt))
            out.nodes[nt].data[k] = best.nodes[nt].data[k]
    for c in cets:
        for k in ref.edges[c].data.keys():
            out.edges[c].data[k] = torch.cat([g.edges[c].data[k] for g in graphs], 0)
    return out
