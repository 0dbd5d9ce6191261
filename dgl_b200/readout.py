"""Graph-level readout on batched graphs.

Every function here is one of three segment operations over the per-graph row ranges of a batched graph's node (or edge)
frame: a segment REDUCTION (sum / mean / max, optionally weighted), a segment SOFTMAX, or a segment TOP-K; ``broadcast_*``
is the inverse mapping (graph rows repeated per member). The node and edge flavours differ only in which frame and which
``batch_num_*`` vector they read, so they are generated from one implementation per operation.
API parity: /root/reference/python/dgl/readout.py:21-640 (readout_nodes / readout_edges and the named variants)."""
from __future__ import annotations

import torch

from .base import DGLError
from .ops import segment_reduce, segment_softmax

__all__ = ["readout_nodes", "readout_edges", "sum_nodes", "sum_edges", "mean_nodes", "mean_edges", "max_nodes",
           "max_edges", "softmax_nodes", "softmax_edges", "broadcast_nodes", "broadcast_edges", "topk_nodes",
           "topk_edges"]


class _Domain:
    """Which frame of the batched graph an operation reads: ("nodes", "ntype") or ("edges", "etype")."""

    def __init__(self, what, type_kw):
        self.what, self.type_kw = what, type_kw

    def unpack(self, graph, kwargs):
        typ = kwargs.pop(self.type_kw, None)
        if kwargs:
            raise TypeError(f"unexpected keyword argument(s) {sorted(kwargs)}; the type selector is `{self.type_kw}`")
        frame = getattr(graph, self.what)[typ].data
        sizes = graph.batch_num_nodes(typ) if self.what == "nodes" else graph.batch_num_edges(typ)
        return frame, sizes


_NODES, _EDGES = _Domain("nodes", "ntype"), _Domain("edges", "etype")


def _readout(domain):
    def readout(graph, feat, weight=None, *, op="sum", **type_kw):
        frame, sizes = domain.unpack(graph, type_kw)
        x = frame[feat] if weight is None else frame[feat] * frame[weight]
        return segment_reduce(sizes, x, op)
    readout.__name__ = f"readout_{domain.what}"
    readout.__doc__ = (f"Per-graph ``op`` (sum / mean / max / min) of the {domain.what[:-1]} feature ``feat`` "
                       f"(times ``weight`` if given); ``{domain.type_kw}`` selects the type on heterographs.")
    return readout


def _named(domain, op):
    generic = _readout(domain)

    def fn(graph, feat, weight=None, **type_kw):
        return generic(graph, feat, weight, op=op, **type_kw)
    fn.__name__ = f"{op}_{domain.what}"
    fn.__doc__ = f"``readout_{domain.what}(..., op='{op}')``."
    return fn


def _softmax(domain):
    def fn(graph, feat, **type_kw):
        frame, sizes = domain.unpack(graph, type_kw)
        return segment_softmax(sizes, frame[feat])
    fn.__name__ = f"softmax_{domain.what}"
    fn.__doc__ = f"Softmax of ``feat`` over the {domain.what} of each graph of the batch."
    return fn


def _broadcast(domain):
    def fn(graph, graph_feat, **type_kw):
        _, sizes = domain.unpack(graph, type_kw)
        if graph.batch_size == 1 and graph_feat.shape[0] != 1:
            graph_feat = graph_feat.unsqueeze(0)                    # an un-batched feature vector of a single graph
        return torch.repeat_interleave(graph_feat, sizes.long().to(graph_feat.device), dim=0)
    fn.__name__ = f"broadcast_{domain.what}"
    fn.__doc__ = f"Row ``i`` of ``graph_feat`` repeated for every {domain.what[:-1]} of graph ``i``."
    return fn


def _topk(domain):
    def fn(graph, feat, k, *, descending=True, sortby=None, **type_kw):
        frame, sizes = domain.unpack(graph, type_kw)
        x = frame[feat]
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        if x.dim() != 2:
            raise DGLError(f"Only support {domain.what} feature `{feat}` with dimension less than or equal to 2")
        sizes = sizes.long().to(x.device)
        num_graphs, width = sizes.shape[0], x.shape[1]
        longest = int(sizes.max()) if num_graphs else 0
        # dense [graphs, slots, width] view padded with the value that sorts last
        pad = float("-inf") if descending else float("inf")
        member_of = torch.repeat_interleave(torch.arange(num_graphs, device=x.device), sizes)
        slot = torch.arange(x.shape[0], device=x.device) - (torch.cumsum(sizes, 0) - sizes)[member_of]
        dense = x.new_full((num_graphs, max(longest, k), width), pad)
        dense[member_of, slot] = x
        if sortby is None:                                           # every column ranked independently
            values, index = torch.topk(dense, k, dim=1, largest=descending)
        else:                                                        # whole rows ranked by one column
            index = torch.argsort(dense[:, :, sortby], dim=1, descending=descending)[:, :k]
            values = torch.gather(dense, 1, index.unsqueeze(-1).expand(-1, -1, width))
        return values.masked_fill(torch.isinf(values), 0.0), index
    fn.__name__ = f"topk_{domain.what}"
    fn.__doc__ = (f"The ``k`` largest (``descending``) {domain.what[:-1]} rows of every graph, ranked per column or by column "
                  "``sortby``; graphs with fewer members are zero padded. Returns (values [B, k, D], positions inside the graph).")
    return fn


readout_nodes, readout_edges = _readout(_NODES), _readout(_EDGES)
sum_nodes, sum_edges = _named(_NODES, "sum"), _named(_EDGES, "sum")
mean_nodes, mean_edges = _named(_NODES, "mean"), _named(_EDGES, "mean")
max_nodes, max_edges = _named(_NODES, "max"), _named(_EDGES, "max")
softmax_nodes, softmax_edges = _softmax(_NODES), _softmax(_EDGES)
broadcast_nodes, broadcast_edges = _broadcast(_NODES), _broadcast(_EDGES)
topk_nodes, topk_edges = _topk(_NODES), _topk(_EDGES)
