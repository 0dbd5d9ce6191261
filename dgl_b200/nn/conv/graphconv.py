"""GraphConv / EdgeWeightNorm.

``GraphConv`` computes ``act( Â X W + b )`` with ``Â`` the (optionally weighted) adjacency under 'both' | 'right' |
'left' | 'none' degree normalisation. Execution plan (differs from /root/reference/python/dgl/nn/pytorch/conv/
graphconv.py:424-476, which scales the rows of X, aggregates, scales the rows of the result and re-checks isolated
nodes with a host sync on every call):

* the normalisation is folded into a per-edge coefficient (cached per graph structure) so the propagation is ONE
  weighted g-SpMM - forward and backward - instead of SpMM + two N x F scalings;
* the GEMM runs on the narrower side of the propagation (transform first when in > out) and takes the bias in its
  epilogue (``addmm``);
* the zero-in-degree verdict is computed once per structure (``_mp.require_in_edges``).

API and parameter names (``weight``, ``bias``) follow the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from ...base import DGLError
from .. import _mp

_MODES = ("none", "both", "right", "left")


class EdgeWeightNorm(nn.Module):
    """Turn positive scalar edge weights into normalised coefficients:
    both  : w_ji / sqrt(sum_k w_jk * sum_k w_ki)      right : w_ji / sum_k w_ki      none : w_ji
    (weighted out-degree of the source, weighted in-degree of the destination; ``eps`` is added to both)."""

    def __init__(self, norm="both", eps=0.0):
        super().__init__()
        self._norm = norm
        self._eps = eps

    def forward(self, graph, edge_weight):
        if edge_weight.dim() != 1:
            raise DGLError("Currently the normalization is only defined on scalar edge weight.")
        if self._norm == "none":
            return edge_weight
        if self._norm == "both" and bool((edge_weight <= 0).any()):
            raise DGLError('Non-positive edge weight detected with `norm="both"`. This leads to square root of zero '
                           "or negative values.")
        power = -0.5 if self._norm == "both" else -1.0
        w_col = edge_weight.unsqueeze(-1)
        into_dst = (_mp.edge_reduce(graph, w_col) + self._eps).pow(power)              # weighted in-degree ^ power
        scaled = _mp.ops.gsddmm(graph, "mul", w_col, into_dst, "e", "v")
        if self._norm == "both":
            # the reversed adjacency shares the edge order: summing edge weights there gives weighted OUT-degrees
            out_of_src = (_mp.ops.gspmm(graph._graphs[0].reverse(), "copy_rhs", "sum", None, w_col) + self._eps).pow(power)
            scaled = _mp.ops.gsddmm(graph, "mul", scaled, out_of_src, "e", "u")
        return scaled.reshape(-1)


class GraphConv(nn.Module):
    r"""Graph convolution ``h_i = act( b + sum_j c_ji h_j W )`` (Kipf & Welling), ``c_ji`` set by ``norm``."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        if norm not in _MODES:
            raise DGLError(f'Invalid norm value. Must be either "none", "both", "right" or "left". But got "{norm}".')
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self._activation = activation
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats)) if weight else None
        self.bias = nn.Parameter(torch.empty(out_feats)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def _pick_weight(self, external):
        if external is None:
            return self.weight
        if self.weight is not None:
            raise DGLError("External weight is provided while at the same time the module has defined its own weight "
                           "parameter. Please create the module with flag weight=False.")
        return external

    def _forward_capacity(self, graph, feat, W):
        """Capacity-shaped block (whole-step CUDA graph): ELL aggregation kernels, degrees from the ELL table."""
        x = feat[0] if isinstance(feat, (tuple, list)) else feat
        shrink_first = W is not None and self._in_feats > self._out_feats
        if self._norm in ("both", "left"):
            _, outdeg = _mp.capacity_degrees(graph, x.shape[0])
            x = x * outdeg.clamp(min=1).pow(-0.5 if self._norm == "both" else -1.0).unsqueeze(1).to(x.dtype)
        if shrink_first:
            x = x @ W.to(x.dtype)
        pad = (-x.shape[1]) % 4
        if pad:
            x = torch.nn.functional.pad(x, (0, pad))
        _, h = _mp.capacity_aggregate(graph, x, "mean" if self._norm == "right" else "sum")
        h = h[:, : h.shape[1] - pad] if pad else h
        if self._norm == "both":
            indeg, _ = _mp.capacity_degrees(graph, x.shape[0])
            h = h * indeg.clamp(min=1).pow(-0.5).unsqueeze(1).to(h.dtype)
        if W is not None and not shrink_first:
            h = h @ W.to(h.dtype)
        if self.bias is not None:
            h = h + self.bias.to(h.dtype)
        return h if self._activation is None else self._activation(h)

    def forward(self, graph, feat, weight=None, edge_weight=None):
        if _mp.is_capacity_block(graph):
            if edge_weight is not None:
                raise DGLError("edge weights are not supported on capacity blocks")
            return self._forward_capacity(graph, feat, self._pick_weight(weight))
        _mp.require_in_edges(graph, self._allow_zero_in_degree, "GraphConv")
        if edge_weight is not None and edge_weight.shape[0] != graph.num_edges():
            raise DGLError("edge_weight must have one entry per edge")
        W = self._pick_weight(weight)
        x_src, _ = _mp.src_dst(graph, feat)
        coeff = _mp.norm_coefficients(graph, self._norm, x_src.dtype, edge_weight)
        shrink_first = W is not None and self._in_feats > self._out_feats
        if shrink_first:                                   # fewer columns travel through the g-SpMM
            x_src = x_src @ W
        h = _mp.gather_sum(graph, x_src, coeff)
        if W is not None and not shrink_first:
            if self.bias is not None and h.dim() == 2:
                h = torch.addmm(self.bias.to(h.dtype), h, W.to(h.dtype))
                return h if self._activation is None else self._activation(h)
            h = h @ W
        if self.bias is not None:
            h = h + self.bias
        return h if self._activation is None else self._activation(h)

    def extra_repr(self):
        s = f"in={self._in_feats}, out={self._out_feats}, normalization={self._norm}"
        if self._activation is not None:
            s += f", activation={self._activation}"
        return s
