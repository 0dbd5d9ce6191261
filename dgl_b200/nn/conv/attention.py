"""Attention convolutions: GATConv, GATv2Conv, DotGatConv, AGNNConv, EGATConv, EdgeGATConv.

API parity: /root/reference/python/dgl/nn/pytorch/conv/{gatconv,gatv2conv,dotgatconv,agnnconv,egatconv,edgegatconv}.py
(constructor signatures, parameter names; values checked by tests/test_nn_reference_parity.py).

Execution plan shared by the layers (``_HeadedAttention``):

* on CUDA the graph is swapped for its destination-sorted view so scores, softmax and aggregation share one edge order
  (no [E, heads] permutation between the COO-ordered g-SDDMM and the CSC-ordered g-SpMM);
* GAT's additive scores go through the FUSED attention kernels (csrc/cuda/gat_fused.cu: online-softmax forward over CSC
  rows, recomputing backward) - no [E, heads] tensor exists in either direction - whenever no per-edge output is asked;
* otherwise: one g-SDDMM for the logits, the fused edge-softmax kernel, one weighted g-SpMM;
* the isolated-node check is evaluated once per graph structure (``_mp.require_in_edges``), not once per forward.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...base import DGLError
from .. import _mp


def _pair(width):
    return tuple(width) if isinstance(width, (tuple, list)) else (width, width)


def _attention_view(graph, *edge_tensors):
    """(graph to compute on, eid map, edge tensors permuted into that graph's edge order)."""
    if graph.device.type != "cuda" or len(graph.canonical_etypes) != 1:
        return (graph, None) + edge_tensors
    view, emap = graph.dst_sorted_view()
    if emap is None:
        return (view, None) + edge_tensors
    idx = emap.long()
    return (view, emap) + tuple(None if t is None else t[idx] for t in edge_tensors)


def _to_eid_order(per_edge, emap):
    if emap is None:
        return per_edge
    out = torch.empty_like(per_edge)
    out[emap.long()] = per_edge
    return out


class _HeadedAttention(nn.Module):
    """Plumbing common to the multi-head layers: endpoint projections, residual / bias / activation epilogue."""

    _layer_name = "attention layer"

    def _make_projections(self, in_feats, width, bias=False):
        """``fc`` shared by both endpoints, or ``fc_src`` / ``fc_dst`` when the two sides have different widths."""
        self._in_src_feats, self._in_dst_feats = _pair(in_feats)
        if isinstance(in_feats, (tuple, list)):
            self.fc_src = nn.Linear(self._in_src_feats, width, bias=bias)
            self.fc_dst = nn.Linear(self._in_dst_feats, width, bias=bias)
        else:
            self.fc = nn.Linear(self._in_src_feats, width, bias=bias)

    def _make_residual(self, residual, width, bias=False):
        if not residual:
            self.register_buffer("res_fc", None)
        elif self._in_dst_feats != width:
            self.res_fc = nn.Linear(self._in_dst_feats, width, bias=bias)
        else:
            self.res_fc = nn.Identity()

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def _inputs(self, graph, feat, drop=None):
        """(h_src, h_dst) after feature dropout; a block's destination rows are the prefix of its source rows."""
        if isinstance(feat, (tuple, list)):
            h_src, h_dst = feat
            return (h_src, h_dst) if drop is None else (drop(h_src), drop(h_dst))
        h = feat if drop is None else drop(feat)
        return h, (h[: graph.number_of_dst_nodes()] if graph.is_block else h)

    def _project(self, graph, h_src, h_dst, shared_input):
        """[.., heads, out] projections of both endpoints; one GEMM when the endpoints share input and weights."""
        H, D = self._num_heads, self._out_feats
        if hasattr(self, "fc"):
            z_src = self.fc(h_src).view(*h_src.shape[:-1], H, D)
            z_dst = z_src[: graph.number_of_dst_nodes()] if shared_input else self.fc(h_dst).view(*h_dst.shape[:-1], H, D)
            return z_src, z_dst
        return (self.fc_src(h_src).view(*h_src.shape[:-1], H, D), self.fc_dst(h_dst).view(*h_dst.shape[:-1], H, D))

    def _epilogue(self, out, h_dst, bias=None):
        H, D = self._num_heads, self._out_feats
        if self.res_fc is not None:
            out = out + self.res_fc(h_dst).view(*h_dst.shape[:-1], -1, D)
        if bias is not None:
            out = out + bias.view(*((1,) * (out.dim() - 2)), H, D)
        return out if not self.activation else self.activation(out)


# =====================================================================================================================
class GATConv(_HeadedAttention):
    r"""Graph attention (Velickovic et al.): ``e_ij = LeakyReLU(a_l . W h_j + a_r . W h_i)``, softmax over in-edges."""

    _layer_name = "GATConv"

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2, residual=False,
                 activation=None, allow_zero_in_degree=False, bias=True):
        super().__init__()
        self._num_heads, self._out_feats, self._allow_zero_in_degree = num_heads, out_feats, allow_zero_in_degree
        width = out_feats * num_heads
        self._make_projections(in_feats, width)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop, self.attn_drop = nn.Dropout(feat_drop), nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self.has_linear_res = bool(residual) and self._in_dst_feats != width
        self._make_residual(residual, width, bias=bias)            # a linear residual carries the bias itself
        self.has_explicit_bias = bool(bias) and not self.has_linear_res
        if self.has_explicit_bias:
            self.bias = nn.Parameter(torch.empty(width))
        else:
            self.register_buffer("bias", None)
        self.activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for lin in (getattr(self, n) for n in ("fc", "fc_src", "fc_dst") if hasattr(self, n)):
            nn.init.xavier_normal_(lin.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        if self.has_explicit_bias:
            nn.init.zeros_(self.bias)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)
            if self.res_fc.bias is not None:
                nn.init.zeros_(self.res_fc.bias)

    def _scores(self, h_src, h_dst, z_src, z_dst):
        """Attention logit halves ``a_l . (W h_j)`` and ``a_r . (W h_i)`` per head. The attention vectors are folded into the
        projection, ``a . (W h) = (W^T a) . h``, so each half is ONE [N, in] x [in, H] product instead of a broadcast
        multiply and a reduction over the [N, H, D] projections (those two elementwise passes and their backward were
        the largest items of the GAT mini-batch step in the profile)."""
        H, D = self._num_heads, self._out_feats
        lin_src, lin_dst = (self.fc, self.fc) if hasattr(self, "fc") else (self.fc_src, self.fc_dst)
        if lin_src.bias is not None or lin_dst.bias is not None or h_src.dim() != 2:
            return (z_src * self.attn_l).sum(-1), (z_dst * self.attn_r).sum(-1)
        w_l = (lin_src.weight.view(H, D, -1) * self.attn_l.view(H, D, 1)).sum(1)      # [H, in]
        w_r = (lin_dst.weight.view(H, D, -1) * self.attn_r.view(H, D, 1)).sum(1)
        return h_src @ w_l.t().to(h_src.dtype), h_dst @ w_r.t().to(h_dst.dtype)

    def _fused(self, g, z_src, s_src, s_dst):
        """Scores + softmax + aggregation in the fused kernels; None when the shape is not covered."""
        from ...ops import gat_fused
        pad = (-self._out_feats) % (16 // z_src.element_size())   # per-head rows padded to 16 bytes
        s_src, s_dst = s_src.to(z_src.dtype), s_dst.to(z_src.dtype)  # autocast leaves the score reductions in fp32
        if not gat_fused.supported(s_src.unsqueeze(-1), z_src, self._num_heads, self._out_feats + pad):
            return None
        z = F.pad(z_src, (0, pad)) if pad else z_src
        out = gat_fused.fused_gat(g._graphs[0], s_src, s_dst, z, self.leaky_relu.negative_slope)
        return out[..., : self._out_feats] if pad else out

    def _forward_capacity(self, graph, feat):
        """Capacity-shaped block (whole-step CUDA graph). Fast path: the fused attention kernels over a CSC / CSR pair
        built on the device from the ELL table (``_mp.CapacityRelation``). Shapes the kernels do not take, or attention
        dropout in training, use dense masked attention over the ELL table ([rows, fanout, heads] tensors of fixed shape -
        functional, but slow: the gathered tensor is materialised and its backward is an index-add over all slots)."""
        if isinstance(feat, (tuple, list)):
            raise DGLError("capacity blocks take one feature tensor (destinations are the prefix of the sources)")
        nbr, cnt, fanout = graph._ell_cache
        h_src = self.feat_drop(feat)
        h_dst = h_src[: graph.num_dst_nodes()]
        z_src, z_dst = self._project(graph, h_src, h_dst, shared_input=True)
        s_src, s_dst = self._scores(h_src, h_dst, z_src, z_dst)   # [N_src, H], [N_dst_cap, H]
        from ...ops import gat_fused
        pad = (-self._out_feats) % (16 // z_src.element_size())
        s_src, s_dst = s_src.to(z_src.dtype), s_dst.to(z_src.dtype)   # autocast leaves the score reductions in fp32
        if not (self.training and self.attn_drop.p > 0) and \
                gat_fused.supported(s_src.unsqueeze(-1), z_src, self._num_heads, self._out_feats + pad):
            rel = _mp.CapacityRelation(graph, z_src.shape[0])
            z = F.pad(z_src, (0, pad)) if pad else z_src
            out = gat_fused.fused_gat(rel, s_src, s_dst, z, self.leaky_relu.negative_slope)
            out = out[..., : self._out_feats] if pad else out
            return self._epilogue(out, h_dst, self.bias)
        slot = nbr.clamp(min=0).long()                             # [N_dst_cap, fanout]
        live = torch.arange(fanout, device=nbr.device).unsqueeze(0) < cnt.unsqueeze(1)
        score = self.leaky_relu(s_src[slot] + s_dst.unsqueeze(1))  # [N_dst_cap, fanout, H]
        score = score.masked_fill(~live.unsqueeze(-1), float("-inf"))
        attn = torch.softmax(score, dim=1)
        attn = torch.nan_to_num(attn, nan=0.0)                     # rows without neighbours: all slots masked
        attn = self.attn_drop(attn)
        out = (attn.unsqueeze(-1) * z_src[slot]).sum(1)            # [N_dst_cap, H, D]
        return self._epilogue(out, h_dst, self.bias)

    def forward(self, graph, feat, edge_weight=None, get_attention=False):
        if _mp.is_capacity_block(graph):
            if edge_weight is not None or get_attention:
                raise DGLError("edge weights / attention output are not available on capacity blocks")
            return self._forward_capacity(graph, feat)
        _mp.require_in_edges(graph, self._allow_zero_in_degree, self._layer_name)
        g, emap, edge_weight = _attention_view(graph, edge_weight)
        h_src, h_dst = self._inputs(g, feat, self.feat_drop)
        z_src, z_dst = self._project(g, h_src, h_dst, shared_input=not isinstance(feat, (tuple, list)))
        s_src, s_dst = self._scores(h_src, h_dst, z_src, z_dst)   # [N_src, H], [N_dst, H]
        plain = edge_weight is None and not get_attention and not (self.training and self.attn_drop.p > 0)
        if plain and z_src.is_cuda and z_src.dim() == 3 and len(g.canonical_etypes) == 1:
            out = self._fused(g, z_src, s_src, s_dst)
            if out is not None:
                return self._epilogue(out, h_dst, self.bias)
        logits = self.leaky_relu(_mp.pair_op(g, "add", s_src.unsqueeze(-1), s_dst.unsqueeze(-1)))
        attn = self.attn_drop(_mp.attend(g, logits))               # [E, H, 1]
        if edge_weight is not None:
            attn = attn * edge_weight.reshape(-1, 1, 1)
        # 4-byte accesses on E x D elements are avoided by padding per-head rows to 16 bytes on dense graphs
        pad = 0
        if z_src.is_cuda and z_src.dim() == 3 and g.num_edges() >= 4 * z_src.shape[0]:
            pad = (-self._out_feats) % (16 // z_src.element_size())
        out = _mp.gather_sum(g, F.pad(z_src, (0, pad)) if pad else z_src, attn)
        out = self._epilogue(out[..., : self._out_feats] if pad else out, h_dst, self.bias)
        return (out, _to_eid_order(attn, emap)) if get_attention else out


class GATv2Conv(_HeadedAttention):
    r"""GATv2 (dynamic attention): ``e_ij = a . LeakyReLU(W_l h_j + W_r h_i)``."""

    _layer_name = "GATv2Conv"

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2, residual=False,
                 activation=None, allow_zero_in_degree=False, bias=True, share_weights=False):
        super().__init__()
        self._num_heads, self._out_feats, self._allow_zero_in_degree = num_heads, out_feats, allow_zero_in_degree
        self._in_src_feats, self._in_dst_feats = _pair(in_feats)
        width = out_feats * num_heads
        self.fc_src = nn.Linear(self._in_src_feats, width, bias=bias)
        self.fc_dst = self.fc_src if share_weights else nn.Linear(self._in_dst_feats, width, bias=bias)
        self.attn = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop, self.attn_drop = nn.Dropout(feat_drop), nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self._make_residual(residual, width, bias=bias)
        self.activation, self.share_weights, self.bias = activation, share_weights, bias
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for lin in {self.fc_src, self.fc_dst} | ({self.res_fc} if isinstance(self.res_fc, nn.Linear) else set()):
            nn.init.xavier_normal_(lin.weight, gain=gain)
            if lin.bias is not None:
                nn.init.zeros_(lin.bias)
        nn.init.xavier_normal_(self.attn, gain=gain)

    def forward(self, graph, feat, get_attention=False):
        _mp.require_in_edges(graph, self._allow_zero_in_degree, self._layer_name)
        g, emap = _attention_view(graph)
        h_src, h_dst = self._inputs(g, feat, self.feat_drop)
        H, D = self._num_heads, self._out_feats
        z_src = self.fc_src(h_src).view(-1, H, D)
        if self.share_weights and not isinstance(feat, (tuple, list)):
            z_dst = z_src[: g.number_of_dst_nodes()]
        else:
            z_dst = self.fc_dst(h_dst).view(-1, H, D)
        hidden = self.leaky_relu(_mp.pair_op(g, "add", z_src, z_dst))                    # [E, H, D]
        attn = self.attn_drop(_mp.attend(g, (hidden * self.attn).sum(-1, keepdim=True)))
        out = self._epilogue(_mp.gather_sum(g, z_src, attn), h_dst)
        return (out, _to_eid_order(attn, emap)) if get_attention else out


class DotGatConv(_HeadedAttention):
    r"""Scaled dot-product attention ``e_ij = (W h_i) . (W h_j) / sqrt(d)`` (g-SDDMM u_dot_v)."""

    _layer_name = "DotGatConv"

    def __init__(self, in_feats, out_feats, num_heads, allow_zero_in_degree=False):
        super().__init__()
        self._num_heads, self._out_feats, self._allow_zero_in_degree = num_heads, out_feats, allow_zero_in_degree
        self._make_projections(in_feats, out_feats * num_heads)

    def forward(self, graph, feat, get_attention=False):
        _mp.require_in_edges(graph, self._allow_zero_in_degree, self._layer_name)
        h_src, h_dst = self._inputs(graph, feat)
        z_src, z_dst = self._project(graph, h_src, h_dst, shared_input=not isinstance(feat, (tuple, list)))
        attn = _mp.attend(graph, _mp.pair_op(graph, "dot", z_src, z_dst) / self._out_feats ** 0.5)
        out = _mp.gather_sum(graph, z_src, attn)
        return (out, attn) if get_attention else out


class AGNNConv(nn.Module):
    r"""Attention-based propagation ``P_ij = softmax_j( beta * cos(h_i, h_j) )``, ``H' = P H`` (no weights)."""

    def __init__(self, init_beta=1.0, learn_beta=True, allow_zero_in_degree=False):
        super().__init__()
        self._allow_zero_in_degree = allow_zero_in_degree
        beta = torch.tensor([float(init_beta)])
        if learn_beta:
            self.beta = nn.Parameter(beta)
        else:
            self.register_buffer("beta", beta)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat):
        _mp.require_in_edges(graph, self._allow_zero_in_degree, "AGNNConv")
        x_src, x_dst = _mp.src_dst(graph, feat)
        unit_src = F.normalize(x_src, p=2, dim=-1)
        unit_dst = unit_src if x_dst is x_src else F.normalize(x_dst, p=2, dim=-1)
        cos = _mp.pair_op(graph, "dot", unit_src, unit_dst)
        return _mp.gather_sum(graph, x_src, _mp.attend(graph, self.beta * cos))


class EGATConv(nn.Module):
    r"""GAT with edge features: returns new node features [N, H, out_node] and edge features [E, H, out_edge]."""

    def __init__(self, in_node_feats, in_edge_feats, out_node_feats, out_edge_feats, num_heads, bias=True):
        super().__init__()
        self._num_heads, self._out_node_feats, self._out_edge_feats = num_heads, out_node_feats, out_edge_feats
        self._in_src_node_feats, self._in_dst_node_feats = _pair(in_node_feats)
        dst_width = self._in_dst_node_feats if isinstance(in_node_feats, (tuple, list)) else self._in_src_node_feats
        self.fc_node_src = nn.Linear(self._in_src_node_feats, out_node_feats * num_heads, bias=False)
        self.fc_ni = nn.Linear(self._in_src_node_feats, out_edge_feats * num_heads, bias=False)
        self.fc_nj = nn.Linear(dst_width, out_edge_feats * num_heads, bias=False)
        self.fc_fij = nn.Linear(in_edge_feats, out_edge_feats * num_heads, bias=False)
        self.attn = nn.Parameter(torch.empty(1, num_heads, out_edge_feats))
        if bias:
            self.bias = nn.Parameter(torch.empty(num_heads * out_edge_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for lin in (self.fc_node_src, self.fc_ni, self.fc_fij, self.fc_nj):
            nn.init.xavier_normal_(lin.weight, gain=gain)
        nn.init.xavier_normal_(self.attn, gain=gain)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph, nfeats, efeats, get_attention=False):
        _mp.require_in_edges(graph, False, "EGATConv")
        x_src, x_dst = _mp.src_dst(graph, nfeats)
        H = self._num_heads
        f = _mp.pair_op(graph, "add", self.fc_ni(x_src), self.fc_nj(x_dst)) + self.fc_fij(efeats)
        if self.bias is not None:
            f = f + self.bias
        f_out = F.leaky_relu(f).view(-1, H, self._out_edge_feats)
        attn = _mp.attend(graph, (f_out * self.attn).sum(-1, keepdim=True))
        h_out = _mp.gather_sum(graph, self.fc_node_src(x_src).view(-1, H, self._out_node_feats), attn)
        return (h_out, f_out, attn) if get_attention else (h_out, f_out)


class EdgeGATConv(_HeadedAttention):
    r"""GAT whose logits and messages both include a projection of the edge features:
    ``e_ij = LeakyReLU(a_l . W h_j + a_r . W h_i + a_e . W_e f_ij)``, message ``alpha_ij (W h_j + W_e f_ij)``."""

    _layer_name = "EdgeGATConv"

    def __init__(self, in_feats, edge_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2,
                 residual=True, activation=None, allow_zero_in_degree=False, bias=True):
        super().__init__()
        self._num_heads, self._out_feats, self._allow_zero_in_degree = num_heads, out_feats, allow_zero_in_degree
        width = out_feats * num_heads
        self._make_projections(in_feats, width)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.feat_drop, self.attn_drop = nn.Dropout(feat_drop), nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self.fc_edge = nn.Linear(edge_feats, width, bias=False)
        self.attn_edge = nn.Parameter(torch.empty(1, num_heads, out_feats))
        if bias:
            self.bias = nn.Parameter(torch.empty(width))
        else:
            self.register_buffer("bias", None)
        self._make_residual(residual, width)
        self.activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        lins = [getattr(self, n) for n in ("fc", "fc_src", "fc_dst") if hasattr(self, n)] + [self.fc_edge]
        if isinstance(self.res_fc, nn.Linear):
            lins.append(self.res_fc)
        for lin in lins:
            nn.init.xavier_normal_(lin.weight, gain=gain)
        for a in (self.attn_l, self.attn_r, self.attn_edge):
            nn.init.xavier_normal_(a, gain=gain)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, graph, feat, edge_feat, get_attention=False):
        _mp.require_in_edges(graph, self._allow_zero_in_degree, self._layer_name)
        H, D = self._num_heads, self._out_feats
        h_src, h_dst = self._inputs(graph, feat, self.feat_drop)
        z_src, z_dst = self._project(graph, h_src, h_dst, shared_input=not isinstance(feat, (tuple, list)))
        z_edge = self.fc_edge(edge_feat).view(*edge_feat.shape[:-1], H, D)
        s_src = (z_src * self.attn_l).sum(-1, keepdim=True)
        s_dst = (z_dst * self.attn_r).sum(-1, keepdim=True)
        s_edge = (z_edge * self.attn_edge).sum(-1, keepdim=True)
        attn = self.attn_drop(_mp.attend(graph, self.leaky_relu(_mp.pair_op(graph, "add", s_src, s_dst) + s_edge)))
        out = _mp.gather_sum(graph, z_src, attn) + _mp.edge_reduce(graph, z_edge * attn, "sum")
        out = self._epilogue(out, h_dst, self.bias)
        return (out, attn) if get_attention else out
