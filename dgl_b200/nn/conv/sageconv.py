"""SAGEConv (parity: /root/reference/python/dgl/nn/pytorch/conv/sageconv.py:13-294)."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from ... import function as fn
from ...base import DGLError
from ...utils import check_eq_shape, expand_as_pair


class SAGEConv(nn.Module):
    r"""GraphSAGE layer: ``h_i' = W_self h_i + W_neigh * AGG_{j in N(i)} h_j`` with AGG in
    {'mean', 'gcn', 'pool', 'lstm'} ('gcn' folds the node itself into a degree-normalised sum and has no W_self).

    Execution: the bf16 'mean' case runs as one fused tcgen05 kernel (``nn/fused_sage.py``); everything else calls the
    g-SpMM operators directly — one reduction kernel per layer — and applies ``W_neigh`` on whichever side of the
    aggregation is narrower (before it when ``in > out``), which is what bounds the HBM traffic of the gather.
    Parameter names follow the reference layer (sageconv.py:13-294) so checkpoints are interchangeable."""

    AGGREGATORS = ("mean", "gcn", "pool", "lstm")

    def __init__(self, in_feats, out_feats, aggregator_type, feat_drop=0.0, bias=True, norm=None, activation=None):
        super().__init__()
        if aggregator_type not in self.AGGREGATORS:
            raise DGLError(f"Invalid aggregator_type. Must be one of {set(self.AGGREGATORS)}. "
                           f"But got {aggregator_type!r} instead.")
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats, self._aggre_type = out_feats, aggregator_type
        self.norm, self.activation = norm, activation
        self.feat_drop = nn.Dropout(feat_drop)
        d_src = self._in_src_feats
        if aggregator_type == "pool":
            self.fc_pool = nn.Linear(d_src, d_src)
        elif aggregator_type == "lstm":
            self.lstm = nn.LSTM(d_src, d_src, batch_first=True)
        self.fc_neigh = nn.Linear(d_src, out_feats, bias=False)
        if aggregator_type == "gcn":
            # no separate self projection: the bias lives on the layer
            if bias:
                self.bias = nn.Parameter(torch.zeros(out_feats))
            else:
                self.register_buffer("bias", None)
        else:
            self.fc_self = nn.Linear(self._in_dst_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for name in ("fc_pool", "fc_self", "fc_neigh"):
            if hasattr(self, name):
                nn.init.xavier_uniform_(getattr(self, name).weight, gain=gain)
        if hasattr(self, "lstm"):
            self.lstm.reset_parameters()

    # ---- aggregation strategies: (graph, source feats, dest feats, edge weight) -> W_neigh-projected neighbourhood ----
    @staticmethod
    def _reduce(graph, x, edge_weight, how):
        """One g-SpMM: copy_u (or u_mul_e with per-edge weights) reduced with ``how``."""
        from ... import ops
        if graph.num_edges() == 0:
            return x.new_zeros((graph.num_dst_nodes(),) + tuple(x.shape[1:]))
        if edge_weight is None:
            return getattr(ops, f"copy_u_{how}")(graph, x)
        w = edge_weight.view(-1, *([1] * (x.dim() - 1)))
        return getattr(ops, f"u_mul_e_{how}")(graph, x, w)

    def _project_around(self, aggregate, feat_src):
        """Apply ``W_neigh`` before the aggregation when that shrinks the gathered rows, after it otherwise
        (both orders are the same linear map for sum / mean)."""
        if self._in_src_feats > self._out_feats:
            return aggregate(self.fc_neigh(feat_src)), True
        return self.fc_neigh(aggregate(feat_src)), False

    def _agg_mean(self, graph, feat_src, feat_dst, edge_weight):
        out, _ = self._project_around(lambda x: self._reduce(graph, x, edge_weight, "mean"), feat_src)
        return out

    def _agg_gcn(self, graph, feat_src, feat_dst, edge_weight):
        # (sum over neighbours + the node itself) / (in-degree + 1); W_neigh commutes with the whole expression
        deg = graph.in_degrees().to(feat_dst).unsqueeze(-1) + 1
        if self._in_src_feats > self._out_feats:
            ps = self.fc_neigh(feat_src)
            shares_rows = feat_dst is feat_src or (feat_dst.data_ptr() == feat_src.data_ptr() and feat_dst.numel() > 0)
            pd = ps[: feat_dst.shape[0]] if shares_rows else self.fc_neigh(feat_dst)
            return (self._reduce(graph, ps, edge_weight, "sum") + pd) / deg
        return self.fc_neigh((self._reduce(graph, feat_src, edge_weight, "sum") + feat_dst) / deg)

    def _agg_pool(self, graph, feat_src, feat_dst, edge_weight):
        return self.fc_neigh(self._reduce(graph, F.relu(self.fc_pool(feat_src)), edge_weight, "max"))

    def _agg_lstm(self, graph, feat_src, feat_dst, edge_weight):
        # order-dependent aggregator: needs the per-node mailbox, i.e. the degree-bucketed UDF path
        def run_lstm(nodes):
            seq = nodes.mailbox["m"]
            zeros = seq.new_zeros((1, seq.shape[0], self._in_src_feats))
            _, (last, _) = self.lstm(seq, (zeros, zeros))
            return {"neigh": last.squeeze(0)}

        with graph.local_scope():
            graph.srcdata["h"] = feat_src
            if edge_weight is None:
                message = fn.copy_u("h", "m")
            else:
                graph.edata["_w"] = edge_weight
                message = fn.u_mul_e("h", "_w", "m")
            if graph.num_edges() == 0:
                neigh = feat_src.new_zeros(graph.num_dst_nodes(), self._in_src_feats)
            else:
                graph.update_all(message, run_lstm)
                neigh = graph.dstdata["neigh"]
        return self.fc_neigh(neigh)

    def _try_fused(self, graph, feat, edge_weight):
        if self._aggre_type != "mean" or self.norm is not None:
            return None
        from .. import fused_sage
        if not fused_sage.usable(graph, feat, self._out_feats, edge_weight):
            return None
        # one tcgen05 kernel: gather + mean + concat + [W_self | W_neigh] GEMM + bias (+ ReLU)
        relu = self.activation is F.relu or isinstance(self.activation, nn.ReLU)
        out = fused_sage.fused_sage_mean(graph, self.feat_drop(feat), self.fc_self.weight, self.fc_neigh.weight,
                                         self.fc_self.bias, relu=relu)
        if out is not None and self.activation is not None and not relu:
            out = self.activation(out)
        return out

    def forward(self, graph, feat, edge_weight=None):
        fused = self._try_fused(graph, feat, edge_weight)
        if fused is not None:
            return fused
        from .. import _mp
        if _mp.is_capacity_block(graph):
            # shapes the tcgen05 layer does not take (e.g. > 256 columns): ELL aggregation kernels + plain linear layers
            if self._aggre_type not in ("mean", "gcn") or isinstance(feat, tuple):
                raise DGLError("capacity blocks support the 'mean' and 'gcn' aggregators on a single feature tensor")
            x = self.feat_drop(feat)
            pad = (-x.shape[1]) % 4
            own, pooled = _mp.capacity_aggregate(graph, torch.nn.functional.pad(x, (0, pad)) if pad else x, "mean")
            own, pooled = (own[:, : x.shape[1]], pooled[:, : x.shape[1]]) if pad else (own, pooled)
            if self._aggre_type == "gcn":
                nbr, cnt, _ = graph._ell_cache
                c = cnt.clamp(min=0).unsqueeze(1).to(pooled.dtype)
                out = self.fc_neigh((pooled * c + own) / (c + 1))
                out = out if self.bias is None else out + self.bias
            else:
                out = self.fc_self(own) + self.fc_neigh(pooled)
            out = out if self.activation is None else self.activation(out)
            return out if self.norm is None else self.norm(out)
        if edge_weight is not None and edge_weight.shape[0] != graph.num_edges():
            raise DGLError(f"edge_weight has {edge_weight.shape[0]} rows but the graph has {graph.num_edges()} edges")
        if isinstance(feat, tuple):
            feat_src, feat_dst = self.feat_drop(feat[0]), self.feat_drop(feat[1])
        else:
            feat_src = feat_dst = self.feat_drop(feat)
            if graph.is_block:
                feat_dst = feat_src[: graph.number_of_dst_nodes()]
        if self._aggre_type == "gcn":
            check_eq_shape(feat)
        neigh = getattr(self, "_agg_" + self._aggre_type)(graph, feat_src, feat_dst, edge_weight)
        if self._aggre_type == "gcn":
            out = neigh if self.bias is None else neigh + self.bias
        else:
            out = self.fc_self(feat_dst) + neigh
        if self.activation is not None:
            out = self.activation(out)
        if self.norm is not None:
            out = self.norm(out)
        return out
