"""Graph-level readout on batched graphs, built on the segment-reduce kernel.

Parity: /root/reference/python/dgl/readout.py ({sum,mean,max,softmax,topk}_{nodes,edges},
readout_nodes/edges, broadcast_nodes/edges).
"""
from __future__ import annotations

import torch

from .base import DGLError
from .ops import segment_reduce, segment_softmax

__all__ = ["readout_nodes", "readout_edges", "sum_nodes", "sum_edges", "mean_nodes", "mean_edges", "max_nodes",
           "max_edges", "softmax_nodes", "softmax_edges", "broadcast_nodes", "broadcast_edges", "topk_nodes",
           "topk_edges"]


def readout_nodes(graph, feat, weight=None, *, op="sum", ntype=None):
    x = graph.nodes[ntype].data[feat]
    if weight is not None:
        x = x * graph.nodes[ntype].data[weight]
    return segment_reduce(graph.batch_num_nodes(ntype), x, op)


def readout_edges(graph, feat, weight=None, *, op="sum", etype=None):
    x = graph.edges[etype].data[feat]
    if weight is not None:
        x = x * graph.edges[etype].data[weight]
    return segment_reduce(graph.batch_num_edges(etype), x, op)


def sum_nodes(graph, feat, weight=None, *, ntype=None):
    return readout_nodes(graph, feat, weight, ntype=ntype, op="sum")


def sum_edges(graph, feat, weight=None, *, etype=None):
    return readout_edges(graph, feat, weight, etype=etype, op="sum")


def mean_nodes(graph, feat, weight=None, *, ntype=None):
    return readout_nodes(graph, feat, weight, ntype=ntype, op="mean")


def mean_edges(graph, feat, weight=None, *, etype=None):
    return readout_edges(graph, feat, weight, etype=etype, op="mean")


def max_nodes(graph, feat, weight=None, *, ntype=None):
    return readout_nodes(graph, feat, weight, ntype=ntype, op="max")


def max_edges(graph, feat, weight=None, *, etype=None):
    return readout_edges(graph, feat, weight, etype=etype, op="max")


def softmax_nodes(graph, feat, *, ntype=None):
    return segment_softmax(graph.batch_num_nodes(ntype), graph.nodes[ntype].data[feat])


def softmax_edges(graph, feat, *, etype=None):
    return segment_softmax(graph.batch_num_edges(etype), graph.edges[etype].data[feat])


def broadcast_nodes(graph, graph_feat, *, ntype=None):
    if graph_feat.shape[0] != graph.batch_size and graph.batch_size == 1:
        graph_feat = graph_feat.unsqueeze(0)
    return torch.repeat_interleave(graph_feat, graph.batch_num_nodes(ntype).long().to(graph_feat.device), dim=0)


def broadcast_edges(graph, graph_feat, *, etype=None):
    if graph_feat.shape[0] != graph.batch_size and graph.batch_size == 1:
        graph_feat = graph_feat.unsqueeze(0)
    return torch.repeat_interleave(graph_feat, graph.batch_num_edges(etype).long().to(graph_feat.device), dim=0)


def _topk_on(graph, typestr, feat, k, descending, sortby, ntype_or_etype):
    if typestr == "nodes":
        data = graph.nodes[ntype_or_etype].data
        batch_num = graph.batch_num_nodes(ntype_or_etype)
    else:
        data = graph.edges[ntype_or_etype].data
        batch_num = graph.batch_num_edges(ntype_or_etype)
    x = data[feat]
    if x.dim() == 1:
        x = x.unsqueeze(-1)
    if x.dim() > 2:
        raise DGLError("Only support {} feature `{}` with dimension less than or equal to 2".format(typestr, feat))
    B = batch_num.shape[0]
    hidden = x.shape[-1]
    bn = batch_num.long()
    maxn = int(bn.max()) if B else 0
    fill = float("-inf") if descending else float("inf")
    # pad to [B, maxn, D]
    idx_in_graph = torch.arange(x.shape[0], device=x.device) - torch.repeat_interleave(torch.cumsum(bn, 0) - bn, bn)
    gid = torch.repeat_interleave(torch.arange(B, device=x.device), bn)
    padded = x.new_full((B, max(maxn, k), hidden), fill)
    padded[gid, idx_in_graph] = x
    if sortby is not None:
        keys = padded[:, :, sortby]
        order = torch.argsort(keys, dim=1, descending=descending)
        topk_idx = order[:, :k]
        out = torch.gather(padded, 1, topk_idx.unsqueeze(-1).expand(-1, -1, hidden))
        ret_idx = topk_idx
    else:
        out, ret_idx = torch.topk(padded, k, dim=1, largest=descending)
    out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
    return out, ret_idx


def topk_nodes(graph, feat, k, *, descending=True, sortby=None, ntype=None):
    return _topk_on(graph, "nodes", feat, k, descending, sortby, ntype)


def topk_edges(graph, feat, k, *, descending=True, sortby=None, etype=None):
    return _topk_on(graph, "edges", feat, k, descending, sortby, etype)
