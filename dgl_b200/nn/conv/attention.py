"""Attention-style graph convolutions: GATConv, GATv2Conv, DotGatConv, AGNNConv, EGATConv, EdgeGATConv.

Parity: /root/reference/python/dgl/nn/pytorch/conv/{gatconv,gatv2conv,dotgatconv,agnnconv,egatconv,edgegatconv}.py.
All of them lower to g-SDDMM (u_add_v / u_dot_v) + the fused edge-softmax kernel + g-SpMM (u_mul_e_sum).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import function as fn
from ...base import DGLError
from ...ops import edge_softmax
from ...utils import expand_as_pair


def _zero_in_degree_check(graph, allow):
    if not allow and bool((graph.in_degrees() == 0).any()):
        raise DGLError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid. This is "
                       "harmful for some applications, causing silent performance regression. Adding self-loop on "
                       "the input graph by calling `g = dgl.add_self_loop(g)` will resolve the issue. Setting "
                       "``allow_zero_in_degree`` to be `True` when constructing this module will suppress the check "
                       "and let the code run.")


def _attention_view(graph, edge_weight=None):
    """(graph to compute on, eid map, edge weight in that order): CUDA graphs with one relation are swapped for their
    destination-sorted view so that scores / softmax / aggregation share one edge order (no [E, heads] permutations)."""
    if graph.device.type != "cuda" or len(graph.canonical_etypes) != 1:
        return graph, None, edge_weight
    view, emap = graph.dst_sorted_view()
    if emap is not None and edge_weight is not None:
        edge_weight = edge_weight[emap.long()]
    return view, emap, edge_weight


def _attention_to_eid_order(a, emap):
    if emap is None:
        return a
    out = torch.empty_like(a)
    out[emap.long()] = a
    return out



class GATConv(nn.Module):
    r"""Graph attention layer (Velickovic et al.): e_ij = LeakyReLU(a^T [W h_i || W h_j]) (reference gatconv.py:14-365)."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2, residual=False,
                 activation=None, allow_zero_in_degree=False, bias=True):
        super().__init__()
        self._num_heads = num_heads
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        if isinstance(in_feats, tuple):
            self.fc_src = nn.Linear(self._in_src_feats, out_feats * num_heads, bias=False)
            self.fc_dst = nn.Linear(self._in_dst_feats, out_feats * num_heads, bias=False)
        else:
            self.fc = nn.Linear(self._in_src_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.attn_r = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self.has_linear_res = False
        self.has_explicit_bias = False
        if residual:
            if self._in_dst_feats != out_feats * num_heads:
                self.res_fc = nn.Linear(self._in_dst_feats, num_heads * out_feats, bias=bias)
                self.has_linear_res = True
            else:
                self.res_fc = nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        if bias and not self.has_linear_res:
            self.bias = nn.Parameter(torch.FloatTensor(size=(num_heads * out_feats,)))
            self.has_explicit_bias = True
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()
        self.activation = activation

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        if hasattr(self, "fc"):
            nn.init.xavier_normal_(self.fc.weight, gain=gain)
        else:
            nn.init.xavier_normal_(self.fc_src.weight, gain=gain)
            nn.init.xavier_normal_(self.fc_dst.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        if self.has_explicit_bias:
            nn.init.constant_(self.bias, 0)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)
            if self.res_fc.bias is not None:
                nn.init.constant_(self.res_fc.bias, 0)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, edge_weight=None, get_attention=False):
        _zero_in_degree_check(graph, self._allow_zero_in_degree)
        graph, emap, edge_weight = _attention_view(graph, edge_weight)
        with graph.local_scope():
            if isinstance(feat, tuple):
                src_prefix_shape = feat[0].shape[:-1]
                dst_prefix_shape = feat[1].shape[:-1]
                h_src = self.feat_drop(feat[0])
                h_dst = self.feat_drop(feat[1])
                if not hasattr(self, "fc_src"):
                    feat_src = self.fc(h_src).view(*src_prefix_shape, self._num_heads, self._out_feats)
                    feat_dst = self.fc(h_dst).view(*dst_prefix_shape, self._num_heads, self._out_feats)
                else:
                    feat_src = self.fc_src(h_src).view(*src_prefix_shape, self._num_heads, self._out_feats)
                    feat_dst = self.fc_dst(h_dst).view(*dst_prefix_shape, self._num_heads, self._out_feats)
            else:
                src_prefix_shape = dst_prefix_shape = feat.shape[:-1]
                h_src = h_dst = self.feat_drop(feat)
                feat_src = feat_dst = self.fc(h_src).view(*src_prefix_shape, self._num_heads, self._out_feats)
                if graph.is_block:
                    feat_dst = feat_src[: graph.number_of_dst_nodes()]
                    h_dst = h_dst[: graph.number_of_dst_nodes()]
                    dst_prefix_shape = (graph.number_of_dst_nodes(),) + dst_prefix_shape[1:]
            el = (feat_src * self.attn_l).sum(dim=-1).unsqueeze(-1)
            er = (feat_dst * self.attn_r).sum(dim=-1).unsqueeze(-1)
            # per-head rows that are not a multiple of 16 bytes (e.g. 41 classes) would force 4-byte accesses on E x D
            # elements in the aggregation and its backward: zero-pad the (small) node tensor, slice the result
            ft_pad = 0
            if feat_src.is_cuda and feat_src.dim() == 3 and graph.num_edges() >= 4 * feat_src.shape[0]:
                vec = 16 // feat_src.element_size()
                ft_pad = (-self._out_feats) % vec
            if (feat_src.is_cuda and feat_src.dim() == 3 and edge_weight is None and not get_attention
                    and (self.attn_drop.p == 0 or not self.training) and len(graph.canonical_etypes) == 1):
                # whole attention block in two kernels (online-softmax forward, recomputing backward)
                from ...ops import gat_fused
                vec = 16 // feat_src.element_size()
                pad = (-self._out_feats) % vec
                if gat_fused.supported(el, feat_src, self._num_heads, self._out_feats + pad):
                    ftp = F.pad(feat_src, (0, pad)) if pad else feat_src
                    rst = gat_fused.fused_gat(graph._graphs[0], el.squeeze(-1), er.squeeze(-1), ftp,
                                              self.leaky_relu.negative_slope)
                    if pad:
                        rst = rst[..., : self._out_feats]
                    return self._finish(rst, h_dst, dst_prefix_shape)
            if ft_pad:
                feat_src = F.pad(feat_src, (0, ft_pad))
            graph.srcdata.update({"ft": feat_src, "el": el})
            graph.dstdata.update({"er": er})
            graph.apply_edges(fn.u_add_v("el", "er", "e"))
            e = self.leaky_relu(graph.edata.pop("e"))
            graph.edata["a"] = self.attn_drop(edge_softmax(graph, e))
            if edge_weight is not None:
                graph.edata["a"] = graph.edata["a"] * edge_weight.tile(1, self._num_heads, 1).transpose(0, 2)
            graph.update_all(fn.u_mul_e("ft", "a", "m"), fn.sum("m", "ft"))
            rst = graph.dstdata["ft"]
            if ft_pad:
                rst = rst[..., : self._out_feats]
            rst = self._finish(rst, h_dst, dst_prefix_shape)
            if get_attention:
                return rst, _attention_to_eid_order(graph.edata["a"], emap)
            return rst

    def _finish(self, rst, h_dst, dst_prefix_shape):
        if self.res_fc is not None:
            resval = self.res_fc(h_dst).view(*dst_prefix_shape, -1, self._out_feats)
            rst = rst + resval
        if self.has_explicit_bias:
            rst = rst + self.bias.view(*((1,) * len(dst_prefix_shape)), self._num_heads, self._out_feats)
        if self.activation:
            rst = self.activation(rst)
        return rst


class GATv2Conv(nn.Module):
    r"""GATv2: e_ij = a^T LeakyReLU(W_l h_i + W_r h_j) (reference gatv2conv.py)."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2, residual=False,
                 activation=None, allow_zero_in_degree=False, bias=True, share_weights=False):
        super().__init__()
        self._num_heads = num_heads
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self.fc_src = nn.Linear(self._in_src_feats, out_feats * num_heads, bias=bias)
        if share_weights:
            self.fc_dst = self.fc_src
        else:
            self.fc_dst = nn.Linear(self._in_dst_feats, out_feats * num_heads, bias=bias)
        self.attn = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        if residual:
            if self._in_dst_feats != out_feats * num_heads:
                self.res_fc = nn.Linear(self._in_dst_feats, num_heads * out_feats, bias=bias)
            else:
                self.res_fc = nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        self.activation = activation
        self.share_weights = share_weights
        self.bias = bias
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc_src.weight, gain=gain)
        if self.bias:
            nn.init.constant_(self.fc_src.bias, 0)
        if not self.share_weights:
            nn.init.xavier_normal_(self.fc_dst.weight, gain=gain)
            if self.bias:
                nn.init.constant_(self.fc_dst.bias, 0)
        nn.init.xavier_normal_(self.attn, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)
            if self.bias:
                nn.init.constant_(self.res_fc.bias, 0)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, get_attention=False):
        _zero_in_degree_check(graph, self._allow_zero_in_degree)
        graph, emap, _ = _attention_view(graph)
        with graph.local_scope():
            if isinstance(feat, tuple):
                h_src = self.feat_drop(feat[0])
                h_dst = self.feat_drop(feat[1])
                feat_src = self.fc_src(h_src).view(-1, self._num_heads, self._out_feats)
                feat_dst = self.fc_dst(h_dst).view(-1, self._num_heads, self._out_feats)
            else:
                h_src = h_dst = self.feat_drop(feat)
                feat_src = self.fc_src(h_src).view(-1, self._num_heads, self._out_feats)
                if self.share_weights:
                    feat_dst = feat_src
                else:
                    feat_dst = self.fc_dst(h_src).view(-1, self._num_heads, self._out_feats)
                if graph.is_block:
                    feat_dst = feat_dst[: graph.number_of_dst_nodes()]
                    h_dst = h_dst[: graph.number_of_dst_nodes()]
            graph.srcdata.update({"el": feat_src})
            graph.dstdata.update({"er": feat_dst})
            graph.apply_edges(fn.u_add_v("el", "er", "e"))
            e = self.leaky_relu(graph.edata.pop("e"))
            e = (e * self.attn).sum(dim=-1).unsqueeze(dim=2)
            graph.edata["a"] = self.attn_drop(edge_softmax(graph, e))
            graph.update_all(fn.u_mul_e("el", "a", "m"), fn.sum("m", "ft"))
            rst = graph.dstdata["ft"]
            if self.res_fc is not None:
                resval = self.res_fc(h_dst).view(h_dst.shape[0], -1, self._out_feats)
                rst = rst + resval
            if self.activation:
                rst = self.activation(rst)
            if get_attention:
                return rst, _attention_to_eid_order(graph.edata["a"], emap)
            return rst


class DotGatConv(nn.Module):
    r"""Dot-product attention: e_ij = (W h_i)^T (W h_j) (reference dotgatconv.py) — g-SDDMM u_dot_v path."""

    def __init__(self, in_feats, out_feats, num_heads, allow_zero_in_degree=False):
        super().__init__()
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self._num_heads = num_heads
        if isinstance(in_feats, tuple):
            self.fc_src = nn.Linear(self._in_src_feats, self._out_feats * self._num_heads, bias=False)
            self.fc_dst = nn.Linear(self._in_dst_feats, self._out_feats * self._num_heads, bias=False)
        else:
            self.fc = nn.Linear(self._in_src_feats, self._out_feats * self._num_heads, bias=False)

    def forward(self, graph, feat, get_attention=False):
        graph = graph.local_var()
        _zero_in_degree_check(graph, self._allow_zero_in_degree)
        if isinstance(feat, tuple):
            h_src, h_dst = feat
            feat_src = self.fc_src(h_src).view(-1, self._num_heads, self._out_feats)
            feat_dst = self.fc_dst(h_dst).view(-1, self._num_heads, self._out_feats)
        else:
            h_src = feat
            feat_src = feat_dst = self.fc(h_src).view(-1, self._num_heads, self._out_feats)
            if graph.is_block:
                feat_dst = feat_src[: graph.number_of_dst_nodes()]
        graph.srcdata.update({"ft": feat_src})
        graph.dstdata.update({"ft": feat_dst})
        graph.apply_edges(fn.u_dot_v("ft", "ft", "a"))
        graph.edata["sa"] = edge_softmax(graph, graph.edata["a"] / self._out_feats ** 0.5)
        graph.update_all(fn.u_mul_e("ft", "sa", "attn"), fn.sum("attn", "agg_u"))
        rst = graph.dstdata["agg_u"]
        if get_attention:
            return rst, graph.edata["sa"]
        return rst


class AGNNConv(nn.Module):
    r"""Attention-based GNN: P_ij = softmax(beta * cos(h_i, h_j)) (reference agnnconv.py)."""

    def __init__(self, init_beta=1.0, learn_beta=True, allow_zero_in_degree=False):
        super().__init__()
        self._allow_zero_in_degree = allow_zero_in_degree
        if learn_beta:
            self.beta = nn.Parameter(torch.Tensor([init_beta]))
        else:
            self.register_buffer("beta", torch.Tensor([init_beta]))

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat):
        with graph.local_scope():
            _zero_in_degree_check(graph, self._allow_zero_in_degree)
            feat_src, feat_dst = expand_as_pair(feat, graph)
            graph.srcdata["h"] = feat_src
            graph.srcdata["norm_h"] = F.normalize(feat_src, p=2, dim=-1)
            if isinstance(feat, tuple) or graph.is_block:
                graph.dstdata["norm_h"] = F.normalize(feat_dst, p=2, dim=-1)
            graph.apply_edges(fn.u_dot_v("norm_h", "norm_h", "cos"))
            cos = graph.edata.pop("cos")
            e = self.beta * cos
            graph.edata["p"] = edge_softmax(graph, e)
            graph.update_all(fn.u_mul_e("h", "p", "m"), fn.sum("m", "h"))
            return graph.dstdata.pop("h")


class EGATConv(nn.Module):
    r"""GAT with edge features, returns new node and edge features (reference egatconv.py)."""

    def __init__(self, in_node_feats, in_edge_feats, out_node_feats, out_edge_feats, num_heads, bias=True):
        super().__init__()
        self._num_heads = num_heads
        self._in_src_node_feats, self._in_dst_node_feats = expand_as_pair(in_node_feats)
        self._out_node_feats = out_node_feats
        self._out_edge_feats = out_edge_feats
        if isinstance(in_node_feats, tuple):
            self.fc_node_src = nn.Linear(self._in_src_node_feats, out_node_feats * num_heads, bias=False)
            self.fc_ni = nn.Linear(self._in_src_node_feats, out_edge_feats * num_heads, bias=False)
            self.fc_nj = nn.Linear(self._in_dst_node_feats, out_edge_feats * num_heads, bias=False)
        else:
            self.fc_node_src = nn.Linear(self._in_src_node_feats, out_node_feats * num_heads, bias=False)
            self.fc_ni = nn.Linear(self._in_src_node_feats, out_edge_feats * num_heads, bias=False)
            self.fc_nj = nn.Linear(self._in_src_node_feats, out_edge_feats * num_heads, bias=False)
        self.fc_fij = nn.Linear(in_edge_feats, out_edge_feats * num_heads, bias=False)
        self.attn = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_edge_feats)))
        if bias:
            self.bias = nn.Parameter(torch.FloatTensor(size=(num_heads * out_edge_feats,)))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for m in (self.fc_node_src, self.fc_ni, self.fc_fij, self.fc_nj):
            nn.init.xavier_normal_(m.weight, gain=gain)
        nn.init.xavier_normal_(self.attn, gain=gain)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)

    def forward(self, graph, nfeats, efeats, get_attention=False):
        with graph.local_scope():
            if bool((graph.in_degrees() == 0).any()):
                raise DGLError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid. "
                               "Adding self-loop on the input graph by calling `g = dgl.add_self_loop(g)` will resolve "
                               "the issue.")
            if isinstance(nfeats, tuple):
                nfeats_src, nfeats_dst = nfeats
            else:
                nfeats_src = nfeats_dst = nfeats
                if graph.is_block:
                    nfeats_dst = nfeats[: graph.number_of_dst_nodes()]
            f_ni = self.fc_ni(nfeats_src)
            f_nj = self.fc_nj(nfeats_dst)
            f_fij = self.fc_fij(efeats)
            graph.srcdata.update({"f_ni": f_ni})
            graph.dstdata.update({"f_nj": f_nj})
            graph.apply_edges(fn.u_add_v("f_ni", "f_nj", "f_tmp"))
            f_out = graph.edata.pop("f_tmp") + f_fij
            if self.bias is not None:
                f_out = f_out + self.bias
            f_out = F.leaky_relu(f_out)
            f_out = f_out.view(-1, self._num_heads, self._out_edge_feats)
            e = (f_out * self.attn).sum(dim=-1).unsqueeze(-1)
            graph.edata["a"] = edge_softmax(graph, e)
            graph.srcdata["h_out"] = self.fc_node_src(nfeats_src).view(-1, self._num_heads, self._out_node_feats)
            graph.update_all(fn.u_mul_e("h_out", "a", "m"), fn.sum("m", "h_out"))
            h_out = graph.dstdata["h_out"].view(-1, self._num_heads, self._out_node_feats)
            if get_attention:
                return h_out, f_out, graph.edata.pop("a")
            return h_out, f_out


class EdgeGATConv(nn.Module):
    r"""GAT whose attention logits and messages include edge features (reference edgegatconv.py)."""

    def __init__(self, in_feats, edge_feats, out_feats, num_heads, feat_drop=0.0, attn_drop=0.0, negative_slope=0.2,
                 residual=True, activation=None, allow_zero_in_degree=False, bias=True):
        super().__init__()
        self._num_heads = num_heads
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        if isinstance(in_feats, tuple):
            self.fc_src = nn.Linear(self._in_src_feats, out_feats * num_heads, bias=False)
            self.fc_dst = nn.Linear(self._in_dst_feats, out_feats * num_heads, bias=False)
        else:
            self.fc = nn.Linear(self._in_src_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.attn_r = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self.fc_edge = nn.Linear(edge_feats, out_feats * num_heads, bias=False)
        self.attn_edge = nn.Parameter(torch.FloatTensor(size=(1, num_heads, out_feats)))
        if bias:
            self.bias = nn.Parameter(torch.FloatTensor(size=(num_heads * out_feats,)))
        else:
            self.register_buffer("bias", None)
        if residual:
            if self._in_dst_feats != out_feats * num_heads:
                self.res_fc = nn.Linear(self._in_dst_feats, num_heads * out_feats, bias=False)
            else:
                self.res_fc = nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        self.reset_parameters()
        self.activation = activation

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        if hasattr(self, "fc"):
            nn.init.xavier_normal_(self.fc.weight, gain=gain)
        else:
            nn.init.xavier_normal_(self.fc_src.weight, gain=gain)
            nn.init.xavier_normal_(self.fc_dst.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)
        nn.init.xavier_normal_(self.fc_edge.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_edge, gain=gain)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, edge_feat, get_attention=False):
        with graph.local_scope():
            _zero_in_degree_check(graph, self._allow_zero_in_degree)
            if isinstance(feat, tuple):
                h_src = self.feat_drop(feat[0])
                h_dst = self.feat_drop(feat[1])
                if not hasattr(self, "fc_src"):
                    feat_src = self.fc(h_src).view(-1, self._num_heads, self._out_feats)
                    feat_dst = self.fc(h_dst).view(-1, self._num_heads, self._out_feats)
                else:
                    feat_src = self.fc_src(h_src).view(-1, self._num_heads, self._out_feats)
                    feat_dst = self.fc_dst(h_dst).view(-1, self._num_heads, self._out_feats)
            else:
                h_src = h_dst = self.feat_drop(feat)
                feat_src = feat_dst = self.fc(h_src).view(-1, self._num_heads, self._out_feats)
                if graph.is_block:
                    feat_dst = feat_src[: graph.number_of_dst_nodes()]
                    h_dst = h_dst[: graph.number_of_dst_nodes()]
            n_edges = edge_feat.shape[:-1]
            feat_edge = self.fc_edge(edge_feat).view(*n_edges, self._num_heads, self._out_feats)
            ee = (feat_edge * self.attn_edge).sum(dim=-1).unsqueeze(-1)
            graph.edata.update({"ft_edge": feat_edge, "ee": ee})
            el = (feat_src * self.attn_l).sum(dim=-1).unsqueeze(-1)
            er = (feat_dst * self.attn_r).sum(dim=-1).unsqueeze(-1)
            graph.srcdata.update({"ft": feat_src, "el": el})
            graph.dstdata.update({"er": er})
            graph.apply_edges(fn.u_add_v("el", "er", "e_tmp"))
            e = self.leaky_relu(graph.edata.pop("e_tmp") + graph.edata["ee"])
            graph.edata["a"] = self.attn_drop(edge_softmax(graph, e))
            # message = a * (ft_src + ft_edge)
            graph.update_all(fn.u_mul_e("ft", "a", "m"), fn.sum("m", "ft_n"))
            graph.edata["m_e"] = graph.edata["ft_edge"] * graph.edata["a"]
            graph.update_all(fn.copy_e("m_e", "m"), fn.sum("m", "ft_e"))
            rst = graph.dstdata["ft_n"] + graph.dstdata["ft_e"]
            if self.res_fc is not None:
                resval = self.res_fc(h_dst).view(h_dst.shape[0], -1, self._out_feats)
                rst = rst + resval
            if self.bias is not None:
                rst = rst + self.bias.view(1, self._num_heads, self._out_feats)
            if self.activation:
                rst = self.activation(rst)
            if get_attention:
                return rst, graph.edata["a"]
            return rst
