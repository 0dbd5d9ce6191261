"""Turning user graph data (edge lists, tensor pairs, scipy matrices, networkx graphs) into index tensors
(parity: /root/reference/python/dgl/utils/data.py). ``convert.graph`` & co. accept the same inputs; these are the
stand-alone helpers the reference exposes under ``dgl.utils``."""
from __future__ import annotations

from collections import namedtuple

import torch

from ..base import DGLError
from .checks import check_valid_idtype

__all__ = ["SparseAdjTuple", "elist2tensor", "scipy2tensor", "networkx2tensor", "networkxbipartite2tensors",
           "graphdata2tensors", "infer_num_nodes", "to_device"]

SparseAdjTuple = namedtuple("SparseAdjTuple", ["format", "arrays"])


def elist2tensor(elist, idtype):
    """[(u, v), ...] -> (src, dst) tensors."""
    pairs = list(elist)
    if not pairs:
        return torch.empty(0, dtype=idtype), torch.empty(0, dtype=idtype)
    t = torch.as_tensor(pairs, dtype=idtype).view(-1, 2)
    return t[:, 0].contiguous(), t[:, 1].contiguous()


def scipy2tensor(spmat, idtype):
    """Structure of a scipy sparse matrix (values are dropped): compressed formats stay compressed."""
    if spmat.format in ("csr", "csc"):
        return SparseAdjTuple(spmat.format, (torch.as_tensor(spmat.indptr, dtype=idtype),
                                             torch.as_tensor(spmat.indices, dtype=idtype),
                                             torch.empty(0, dtype=idtype)))
    coo = spmat.tocoo()
    return SparseAdjTuple("coo", (torch.as_tensor(coo.row, dtype=idtype), torch.as_tensor(coo.col, dtype=idtype)))


def _edges_by_id(nx_graph, attr, endpoints):
    """(src, dst) lists; with ``attr`` the edge attribute of that name gives each edge's position."""
    m = nx_graph.number_of_edges()
    if attr is None:
        pairs = [endpoints(u, v) for u, v in nx_graph.edges()]
        return [p[0] for p in pairs], [p[1] for p in pairs]
    src, dst = [0] * m, [0] * m
    for u, v, data in nx_graph.edges(data=True):
        eid = int(data[attr])
        if not 0 <= eid < m:
            raise DGLError(f"Expect edge IDs to be a non-negative integer smaller than {m:d}, got {eid:d}")
        src[eid], dst[eid] = endpoints(u, v)
    return src, dst


def networkx2tensor(nx_graph, idtype, edge_id_attr_name=None):
    """Nodes are relabelled 0..N-1 in sorted label order; undirected graphs contribute both directions."""
    import networkx as nx
    if not nx_graph.is_directed():
        nx_graph = nx_graph.to_directed()
    nx_graph = nx.convert_node_labels_to_integers(nx_graph, ordering="sorted")
    src, dst = _edges_by_id(nx_graph, edge_id_attr_name, lambda u, v: (u, v))
    return torch.tensor(src, dtype=idtype), torch.tensor(dst, dtype=idtype)


def networkxbipartite2tensors(nx_graph, idtype, top_map, bottom_map, edge_id_attr_name=None):
    """Edges of a networkx bipartite graph (``bipartite`` node attribute 0 / 1) from the top to the bottom set."""
    def endpoints(u, v):
        if u not in top_map:
            raise DGLError(f"Expect the node {u} to have attribute bipartite=0 with edge {(u, v)}")
        if v not in bottom_map:
            raise DGLError(f"Expect the node {v} to have attribute bipartite=1 with edge {(u, v)}")
        return top_map[u], bottom_map[v]

    src, dst = _edges_by_id(nx_graph, edge_id_attr_name, endpoints)
    return torch.tensor(src, dtype=idtype), torch.tensor(dst, dtype=idtype)


def infer_num_nodes(data, bipartite=False):
    """(num_src, num_dst) implied by the data, or ``None`` when the type is unknown."""
    def span(t):
        return int(t.max()) + 1 if len(t) else 0

    if isinstance(data, tuple) and len(data) == 2 and isinstance(data[0], str):
        fmt, arrays = data
        if fmt == "coo":
            nsrc, ndst = span(arrays[0]), span(arrays[1])
        elif fmt == "csr":
            nsrc, ndst = arrays[0].shape[0] - 1, span(arrays[1])
        elif fmt == "csc":
            nsrc, ndst = span(arrays[1]), arrays[0].shape[0] - 1
        else:
            raise ValueError(f"unknown format {fmt}")
    elif isinstance(data, tuple):
        raise TypeError(f"Expected sparse format as a str, but got {type(data[0])}")
    elif hasattr(data, "tocoo") and hasattr(data, "shape"):
        nsrc, ndst = data.shape
    elif hasattr(data, "number_of_nodes"):
        n = data.number_of_nodes()
        if n == 0 or not bipartite:
            nsrc = ndst = n
        else:
            nsrc = sum(1 for _, d in data.nodes(data=True) if d["bipartite"] == 0)
            ndst = n - nsrc
    else:
        return None
    if not bipartite:
        nsrc = ndst = max(nsrc, ndst)
    return int(nsrc), int(ndst)


def graphdata2tensors(data, idtype=None, bipartite=False, infer_node_count=True, **kwargs):
    """Normalise any accepted graph input to ``(SparseAdjTuple, num_src, num_dst)``.

    Accepted: ``(row, col)``, ``('coo'|'csr'|'csc', arrays)``, list of pairs, scipy sparse matrix, networkx graph
    (kwargs ``edge_id_attr_name``, ``top_map``, ``bottom_map`` for the latter)."""
    if isinstance(data, tuple):
        data = SparseAdjTuple(*data) if isinstance(data[0], str) else SparseAdjTuple("coo", data)
    tensor_input = isinstance(data, SparseAdjTuple) and torch.is_tensor(data.arrays[0])
    if idtype is None and not tensor_input:
        idtype = torch.int64
    check_valid_idtype(idtype)
    num_src = num_dst = None
    if isinstance(data, SparseAdjTuple):
        arrays = tuple(a if torch.is_tensor(a) else torch.as_tensor(a) for a in data.arrays)
        if idtype is not None:
            arrays = tuple(a.to(idtype) for a in arrays)
        data = SparseAdjTuple(data.format, arrays)
        if infer_node_count:
            num_src, num_dst = infer_num_nodes(data, bipartite)
    elif isinstance(data, list):
        data = SparseAdjTuple("coo", elist2tensor(data, idtype))
        if infer_node_count:
            num_src, num_dst = infer_num_nodes(data, bipartite)
    elif hasattr(data, "tocoo"):
        if infer_node_count:
            num_src, num_dst = infer_num_nodes(data, bipartite)
        data = scipy2tensor(data, idtype)
    elif hasattr(data, "number_of_nodes") and hasattr(data, "edges"):
        if infer_node_count:
            num_src, num_dst = infer_num_nodes(data, bipartite)
        attr = kwargs.get("edge_id_attr_name")
        if bipartite:
            pair = networkxbipartite2tensors(data, idtype, kwargs.get("top_map"), kwargs.get("bottom_map"), attr)
        else:
            pair = networkx2tensor(data, idtype, attr)
        data = SparseAdjTuple("coo", pair)
    else:
        raise DGLError("Unsupported graph data type:", type(data))
    return data, num_src, num_dst


def to_device(data, device):
    """Move a tensor or a dict of tensors."""
    if isinstance(data, dict):
        return {k: v.to(device) for k, v in data.items()}
    return data.to(device)
