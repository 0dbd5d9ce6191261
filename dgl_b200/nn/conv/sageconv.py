"""SAGEConv (parity: /root/reference/python/dgl/nn/pytorch/conv/sageconv.py:13-294)."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from ... import function as fn
from ...base import DGLError
from ...utils import check_eq_shape, expand_as_pair


class SAGEConv(nn.Module):
    r"""GraphSAGE layer; ``aggregator_type`` in {'mean','gcn','pool','lstm'}.
    h_i = W_self h_i + W_neigh AGG({h_j}). Linear-before-aggregate when in > out
    (``lin_before_mp``, reference sageconv.py:230)."""

    def __init__(self, in_feats, out_feats, aggregator_type, feat_drop=0.0, bias=True, norm=None, activation=None):
        super().__init__()
        valid = {"mean", "gcn", "pool", "lstm"}
        if aggregator_type not in valid:
            raise DGLError(f"Invalid aggregator_type. Must be one of {valid}. But got {aggregator_type!r} instead.")
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats = out_feats
        self._aggre_type = aggregator_type
        self.norm = norm
        self.feat_drop = nn.Dropout(feat_drop)
        self.activation = activation
        if aggregator_type == "pool":
            self.fc_pool = nn.Linear(self._in_src_feats, self._in_src_feats)
        if aggregator_type == "lstm":
            self.lstm = nn.LSTM(self._in_src_feats, self._in_src_feats, batch_first=True)
        self.fc_neigh = nn.Linear(self._in_src_feats, out_feats, bias=False)
        if aggregator_type != "gcn":
            self.fc_self = nn.Linear(self._in_dst_feats, out_feats, bias=bias)
        elif bias:
            self.bias = nn.parameter.Parameter(torch.zeros(self._out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        if self._aggre_type == "pool":
            nn.init.xavier_uniform_(self.fc_pool.weight, gain=gain)
        if self._aggre_type == "lstm":
            self.lstm.reset_parameters()
        if self._aggre_type != "gcn":
            nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def _lstm_reducer(self, nodes):
        m = nodes.mailbox["m"]
        batch_size = m.shape[0]
        h = (m.new_zeros((1, batch_size, self._in_src_feats)), m.new_zeros((1, batch_size, self._in_src_feats)))
        _, (rst, _) = self.lstm(m, h)
        return {"neigh": rst.squeeze(0)}

    def forward(self, graph, feat, edge_weight=None):
        if self._aggre_type == "mean" and self.norm is None:
            from .. import fused_sage
            if fused_sage.usable(graph, feat, self._out_feats, edge_weight):
                # one tcgen05 kernel: gather + mean + concat + [W_self | W_neigh] GEMM + bias (+ ReLU)
                relu = self.activation is torch.nn.functional.relu or isinstance(self.activation, torch.nn.ReLU)
                rst = fused_sage.fused_sage_mean(graph, self.feat_drop(feat), self.fc_self.weight, self.fc_neigh.weight,
                                                 self.fc_self.bias, relu=relu)
                if rst is not None:
                    if self.activation is not None and not relu:
                        rst = self.activation(rst)
                    return rst
        with graph.local_scope():
            if isinstance(feat, tuple):
                feat_src = self.feat_drop(feat[0])
                feat_dst = self.feat_drop(feat[1])
            else:
                feat_src = feat_dst = self.feat_drop(feat)
                if graph.is_block:
                    feat_dst = feat_src[: graph.number_of_dst_nodes()]
            msg_fn = fn.copy_u("h", "m")
            if edge_weight is not None:
                assert edge_weight.shape[0] == graph.num_edges()
                graph.edata["_edge_weight"] = edge_weight
                msg_fn = fn.u_mul_e("h", "_edge_weight", "m")
            h_self = feat_dst
            if graph.num_edges() == 0:
                graph.dstdata["neigh"] = torch.zeros(feat_dst.shape[0], self._in_src_feats).to(feat_dst)
            lin_before_mp = self._in_src_feats > self._out_feats
            if self._aggre_type == "mean":
                graph.srcdata["h"] = self.fc_neigh(feat_src) if lin_before_mp else feat_src
                graph.update_all(msg_fn, fn.mean("m", "neigh"))
                h_neigh = graph.dstdata["neigh"]
                if not lin_before_mp:
                    h_neigh = self.fc_neigh(h_neigh)
            elif self._aggre_type == "gcn":
                check_eq_shape(feat)
                graph.srcdata["h"] = self.fc_neigh(feat_src) if lin_before_mp else feat_src
                if isinstance(feat, tuple):
                    graph.dstdata["h"] = self.fc_neigh(feat_dst) if lin_before_mp else feat_dst
                elif graph.is_block:
                    graph.dstdata["h"] = graph.srcdata["h"][: graph.num_dst_nodes()]
                else:
                    graph.dstdata["h"] = graph.srcdata["h"]
                graph.update_all(msg_fn, fn.sum("m", "neigh"))
                degs = graph.in_degrees().to(feat_dst)
                h_neigh = (graph.dstdata["neigh"] + graph.dstdata["h"]) / (degs.unsqueeze(-1) + 1)
                if not lin_before_mp:
                    h_neigh = self.fc_neigh(h_neigh)
            elif self._aggre_type == "pool":
                graph.srcdata["h"] = F.relu(self.fc_pool(feat_src))
                graph.update_all(msg_fn, fn.max("m", "neigh"))
                h_neigh = self.fc_neigh(graph.dstdata["neigh"])
            elif self._aggre_type == "lstm":
                graph.srcdata["h"] = feat_src
                graph.update_all(msg_fn, self._lstm_reducer)
                h_neigh = self.fc_neigh(graph.dstdata["neigh"])
            if self._aggre_type == "gcn":
                rst = h_neigh
                if self.bias is not None:
                    rst = rst + self.bias
            else:
                rst = self.fc_self(h_self) + h_neigh
            if self.activation is not None:
                rst = self.activation(rst)
            if self.norm is not None:
                rst = self.norm(rst)
            return rst
