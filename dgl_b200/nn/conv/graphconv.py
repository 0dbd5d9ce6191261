"""GraphConv / EdgeWeightNorm (parity: /root/reference/python/dgl/nn/pytorch/conv/graphconv.py:14-476)."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import init

from ... import function as fn
from ...base import DGLError
from ...utils import expand_as_pair


class EdgeWeightNorm(nn.Module):
    """Normalise positive scalar edge weights: 'both' -> c_ji = w/sqrt(d_j d_i); 'right' -> w/d_i;
    'none' -> w (reference graphconv.py:14-132)."""

    def __init__(self, norm="both", eps=0.0):
        super().__init__()
        self._norm = norm
        self._eps = eps

    def forward(self, graph, edge_weight):
        with graph.local_scope():
            if isinstance(graph, tuple) or graph.is_block is None:
                pass
            if edge_weight.dim() != 1:
                raise DGLError("Currently the normalization is only defined on scalar edge weight.")
            if self._norm == "both" and bool((edge_weight <= 0).any()):
                raise DGLError('Non-positive edge weight detected with `norm="both"`. This leads to square root '
                               "of zero or negative values.")
            dev = graph.device
            graph.srcdata["_src_out_w"] = torch.ones(graph.number_of_src_nodes(), device=dev, dtype=edge_weight.dtype)
            graph.dstdata["_dst_in_w"] = torch.ones(graph.number_of_dst_nodes(), device=dev, dtype=edge_weight.dtype)
            graph.edata["_edge_w"] = edge_weight
            if self._norm == "both":
                reversed_g = graph.reverse() if not graph.is_block else None
                if reversed_g is not None:
                    reversed_g.edata["_edge_w"] = edge_weight
                    reversed_g.update_all(fn.copy_e("_edge_w", "m"), fn.sum("m", "out_weight"))
                    degs = reversed_g.dstdata["out_weight"] + self._eps
                else:
                    from ... import ops
                    degs = ops.copy_e_sum(graph._graphs[0].reverse(), edge_weight) + self._eps
                graph.srcdata["_src_out_w"] = torch.pow(degs, -0.5)
            if self._norm in ("both", "right"):
                graph.update_all(fn.copy_e("_edge_w", "m"), fn.sum("m", "in_weight"))
                degs = graph.dstdata["in_weight"] + self._eps
                graph.dstdata["_dst_in_w"] = torch.pow(degs, -0.5 if self._norm == "both" else -1.0)
            graph.apply_edges(lambda e: {"_norm_edge_weights": e.src["_src_out_w"] * e.dst["_dst_in_w"] * e.data["_edge_w"]})
            return graph.edata["_norm_edge_weights"]


class GraphConv(nn.Module):
    r"""Graph convolution h_i = b + sum_j 1/c_ji h_j W (Kipf & Welling).

    ``norm`` in {'both','right','left','none'}; aggregates before or after the weight GEMM
    depending on in_feats > out_feats (reference graphconv.py:447-460)."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        if norm not in ("none", "both", "right", "left"):
            raise DGLError(f'Invalid norm value. Must be either "none", "both", "right" or "left". But got "{norm}".')
        self._in_feats = in_feats
        self._out_feats = out_feats
        self._norm = norm
        self._allow_zero_in_degree = allow_zero_in_degree
        if weight:
            self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats))
        else:
            self.register_parameter("weight", None)
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self._activation = activation

    def reset_parameters(self):
        if self.weight is not None:
            init.xavier_uniform_(self.weight)
        if self.bias is not None:
            init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, weight=None, edge_weight=None):
        with graph.local_scope():
            if not self._allow_zero_in_degree:
                if bool((graph.in_degrees() == 0).any()):
                    raise DGLError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid. "
                                   "This is harmful for some applications, causing silent performance regression. "
                                   "Adding self-loop on the input graph by calling `g = dgl.add_self_loop(g)` will "
                                   "resolve the issue. Setting ``allow_zero_in_degree`` to be `True` when "
                                   "constructing this module will suppress the check and let the code run.")
            aggregate_fn = fn.copy_u("h", "m")
            if edge_weight is not None:
                assert edge_weight.shape[0] == graph.num_edges()
                graph.edata["_edge_weight"] = edge_weight
                aggregate_fn = fn.u_mul_e("h", "_edge_weight", "m")
            feat_src, feat_dst = expand_as_pair(feat, graph)
            if self._norm in ("left", "both"):
                degs = graph.out_degrees().to(feat_src).clamp(min=1)
                norm = torch.pow(degs, -0.5) if self._norm == "both" else 1.0 / degs
                feat_src = feat_src * norm.view(-1, *([1] * (feat_src.dim() - 1)))
            if weight is not None:
                if self.weight is not None:
                    raise DGLError("External weight is provided while at the same time the module has defined its "
                                   "own weight parameter. Please create the module with flag weight=False.")
            else:
                weight = self.weight
            if self._in_feats > self._out_feats:
                if weight is not None:
                    feat_src = torch.matmul(feat_src, weight)
                graph.srcdata["h"] = feat_src
                graph.update_all(aggregate_fn, fn.sum(msg="m", out="h"))
                rst = graph.dstdata["h"]
            else:
                graph.srcdata["h"] = feat_src
                graph.update_all(aggregate_fn, fn.sum(msg="m", out="h"))
                rst = graph.dstdata["h"]
                if weight is not None:
                    rst = torch.matmul(rst, weight)
            if self._norm in ("right", "both"):
                degs = graph.in_degrees().to(feat_dst).clamp(min=1)
                norm = torch.pow(degs, -0.5) if self._norm == "both" else 1.0 / degs
                rst = rst * norm.view(-1, *([1] * (feat_dst.dim() - 1)))
            if self.bias is not None:
                rst = rst + self.bias
            if self._activation is not None:
                rst = self._activation(rst)
            return rst

    def extra_repr(self):
        s = "in={_in_feats}, out={_out_feats}, normalization={_norm}"
        if "_activation" in self.__dict__:
            s += ", activation={_activation}"
        return s.format(**self.__dict__)
