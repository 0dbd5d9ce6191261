"""Module-style transforms (filled in by transforms/module_impl.py)."""
from .module_impl import *  # noqa: F401,F403

import torch as _torch

from ..convert import heterograph as _heterograph


def update_graph_structure(g, data_dict, copy_edata=True):
    """New graph with the node sets and node features of ``g`` and the relations of ``data_dict``; edge features are
    carried over per canonical edge type when ``copy_edata`` (reference transforms/module.py:60)."""
    new_g = _heterograph(data_dict, num_nodes_dict={t: g.num_nodes(t) for t in g.ntypes}, idtype=g.idtype,
                         device=g.device)
    for t in g.ntypes:
        for k, v in g.nodes[t].data.items():
            new_g.nodes[t].data[k] = v
    if copy_edata:
        for c in g.canonical_etypes:
            if c not in new_g.canonical_etypes:
                continue
            for k, v in g.edges[c].data.items():
                if new_g.num_edges(c) == v.shape[0]:
                    new_g.edges[c].data[k] = v
    return new_g


def is_bidirected(g):
    """True iff every edge (u, v) has a matching edge (v, u) with the same multiplicity (reference module.py:1294)."""
    src, dst = g.edges()
    n = g.num_nodes()
    fwd = _torch.sort(src.long() * n + dst.long())[0]
    bwd = _torch.sort(dst.long() * n + src.long())[0]
    return bool(_torch.equal(fwd, bwd))
