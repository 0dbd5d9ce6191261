"""Sum/mean-aggregation and spectral convolutions: GINConv, GINEConv, SGConv, TAGConv, APPNPConv, ChebConv,
GCN2Conv, GatedGraphConv, GatedGCNConv, CFConv, NNConv, GMMConv, TWIRLSConv, EdgeConv, EGNNConv.

Parity: the like-named files under /root/reference/python/dgl/nn/pytorch/conv/. Every layer lowers to the
g-SpMM / g-SDDMM primitives.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import function as fn
from ...base import DGLError
from ...ops import edge_softmax
from ...utils import expand_as_pair


class GINConv(nn.Module):
    r"""h_i' = f((1 + eps) h_i + aggregate_j(e_ji h_j)) (reference ginconv.py)."""

    def __init__(self, apply_func=None, aggregator_type="sum", init_eps=0, learn_eps=False, activation=None):
        super().__init__()
        self.apply_func = apply_func
        self._aggregator_type = aggregator_type
        self.activation = activation
        if aggregator_type not in ("sum", "max", "mean"):
            raise KeyError("Aggregator type {} not recognized.".format(aggregator_type))
        if learn_eps:
            self.eps = nn.Parameter(torch.FloatTensor([init_eps]))
        else:
            self.register_buffer("eps", torch.FloatTensor([init_eps]))

    def forward(self, graph, feat, edge_weight=None):
        _reducer = getattr(fn, self._aggregator_type)
        with graph.local_scope():
            aggregate_fn = fn.copy_u("h", "m")
            if edge_weight is not None:
                assert edge_weight.shape[0] == graph.num_edges()
                graph.edata["_edge_weight"] = edge_weight
                aggregate_fn = fn.u_mul_e("h", "_edge_weight", "m")
            feat_src, feat_dst = expand_as_pair(feat, graph)
            graph.srcdata["h"] = feat_src
            graph.update_all(aggregate_fn, _reducer("m", "neigh"))
            rst = (1 + self.eps) * feat_dst + graph.dstdata["neigh"]
            if self.apply_func is not None:
                rst = self.apply_func(rst)
            if self.activation is not None:
                rst = self.activation(rst)
            return rst


class GINEConv(nn.Module):
    r"""h_i' = f((1+eps) h_i + sum_j ReLU(h_j + e_ji)) (reference gineconv.py)."""

    def __init__(self, apply_func=None, init_eps=0, learn_eps=False):
        super().__init__()
        self.apply_func = apply_func
        if learn_eps:
            self.eps = nn.Parameter(torch.FloatTensor([init_eps]))
        else:
            self.register_buffer("eps", torch.FloatTensor([init_eps]))

    def message(self, edges):
        return {"m": F.relu(edges.src["hn"] + edges.data["he"])}

    def forward(self, graph, node_feat, edge_feat):
        with graph.local_scope():
            feat_src, feat_dst = expand_as_pair(node_feat, graph)
            graph.srcdata["hn"] = feat_src
            graph.edata["he"] = edge_feat
            graph.apply_edges(fn.u_add_e("hn", "he", "m0"))
            graph.edata["m"] = F.relu(graph.edata.pop("m0"))
            graph.update_all(fn.copy_e("m", "m"), fn.sum("m", "neigh"))
            rst = (1 + self.eps) * feat_dst + graph.dstdata["neigh"]
            if self.apply_func is not None:
                rst = self.apply_func(rst)
            return rst


class SGConv(nn.Module):
    r"""Simplified GCN: H = S^K X Theta (reference sgconv.py)."""

    def __init__(self, in_feats, out_feats, k=1, cached=False, bias=True, norm=None, allow_zero_in_degree=False):
        super().__init__()
        self.fc = nn.Linear(in_feats, out_feats, bias=bias)
        self._cached = cached
        self._cached_h = None
        self._k = k
        self.norm = norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.fc.weight)
        if self.fc.bias is not None:
            nn.init.zeros_(self.fc.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, edge_weight=None):
        with graph.local_scope():
            if not self._allow_zero_in_degree and bool((graph.in_degrees() == 0).any()):
                raise DGLError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid. "
                               "Adding self-loop on the input graph by calling `g = dgl.add_self_loop(g)` will "
                               "resolve the issue.")
            msg_func = fn.copy_u("h", "m")
            if edge_weight is not None:
                graph.edata["_edge_weight"] = edge_weight
                msg_func = fn.u_mul_e("h", "_edge_weight", "m")
            if self._cached_h is not None:
                feat = self._cached_h
            else:
                if edge_weight is None:
                    degs = graph.in_degrees().to(feat).clamp(min=1)
                    norm = torch.pow(degs, -0.5).unsqueeze(1)
                for _ in range(self._k):
                    if edge_weight is None:
                        feat = feat * norm
                    graph.ndata["h"] = feat
                    graph.update_all(msg_func, fn.sum("m", "h"))
                    feat = graph.ndata.pop("h")
                    if edge_weight is None:
                        feat = feat * norm
                if self.norm is not None:
                    feat = self.norm(feat)
                if self._cached:
                    self._cached_h = feat
            return self.fc(feat)


class TAGConv(nn.Module):
    r"""Topology adaptive GCN: H = sum_k A^k X Theta_k (reference tagconv.py)."""

    def __init__(self, in_feats, out_feats, k=2, bias=True, activation=None):
        super().__init__()
        self._in_feats, self._out_feats, self._k = in_feats, out_feats, k
        self._activation = activation
        self.lin = nn.Linear(in_feats * (self._k + 1), out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_normal_(self.lin.weight, gain=nn.init.calculate_gain("relu"))

    def forward(self, graph, feat, edge_weight=None):
        with graph.local_scope():
            assert graph.is_homogeneous, "Graph is not homogeneous"
            if edge_weight is None:
                norm = torch.pow(graph.in_degrees().to(feat).clamp(min=1), -0.5)
                norm = norm.view(-1, *([1] * (feat.dim() - 1)))
            msg_func = fn.copy_u("h", "m")
            if edge_weight is not None:
                graph.edata["_edge_weight"] = edge_weight
                msg_func = fn.u_mul_e("h", "_edge_weight", "m")
            fstack = [feat]
            for _ in range(self._k):
                rst = fstack[-1] if edge_weight is not None else fstack[-1] * norm
                graph.ndata["h"] = rst
                graph.update_all(msg_func, fn.sum(msg="m", out="h"))
                rst = graph.ndata["h"]
                if edge_weight is None:
                    rst = rst * norm
                fstack.append(rst)
            rst = self.lin(torch.cat(fstack, dim=-1))
            if self._activation is not None:
                rst = self._activation(rst)
            return rst


class APPNPConv(nn.Module):
    r"""Approximate personalised propagation: H^{t+1} = (1-a) A_hat H^t + a H^0 (reference appnpconv.py)."""

    def __init__(self, k, alpha, edge_drop=0.0):
        super().__init__()
        self._k, self._alpha = k, alpha
        self.edge_drop = nn.Dropout(edge_drop)

    def forward(self, graph, feat, edge_weight=None):
        with graph.local_scope():
            if edge_weight is None:
                src_norm = torch.pow(graph.out_degrees().to(feat).clamp(min=1), -0.5)
                src_norm = src_norm.view(-1, *([1] * (feat.dim() - 1)))
                dst_norm = torch.pow(graph.in_degrees().to(feat).clamp(min=1), -0.5)
                dst_norm = dst_norm.view(-1, *([1] * (feat.dim() - 1)))
            else:
                edge_weight = EdgeWeightNormHelper.both(graph, edge_weight)
            feat_0 = feat
            for _ in range(self._k):
                if edge_weight is None:
                    feat = feat * src_norm
                graph.ndata["h"] = feat
                w = torch.ones(graph.num_edges(), 1, device=feat.device, dtype=feat.dtype) if edge_weight is None \
                    else edge_weight.view(-1, 1).to(feat.dtype)
                graph.edata["w"] = self.edge_drop(w).to(feat.device)
                graph.update_all(fn.u_mul_e("h", "w", "m"), fn.sum("m", "h"))
                feat = graph.ndata.pop("h")
                if edge_weight is None:
                    feat = feat * dst_norm
                feat = (1 - self._alpha) * feat + self._alpha * feat_0
            return feat


class EdgeWeightNormHelper:
    @staticmethod
    def both(graph, w):
        from .graphconv import EdgeWeightNorm
        return EdgeWeightNorm("both")(graph, w)


class ChebConv(nn.Module):
    r"""Chebyshev spectral convolution (reference chebconv.py)."""

    def __init__(self, in_feats, out_feats, k, activation=F.relu, bias=True):
        super().__init__()
        self._k, self._in_feats, self._out_feats = k, in_feats, out_feats
        self.activation = activation
        self.linear = nn.Linear(k * in_feats, out_feats, bias)

    def forward(self, graph, feat, lambda_max=None):
        def unnLaplacian(feat, D_invsqrt, graph):
            graph.ndata["h"] = feat * D_invsqrt
            graph.update_all(fn.copy_u("h", "m"), fn.sum("m", "h"))
            return graph.ndata.pop("h") * D_invsqrt

        with graph.local_scope():
            D_invsqrt = torch.pow(graph.in_degrees().to(feat).clamp(min=1), -0.5).unsqueeze(-1)
            if lambda_max is None:
                from ...transforms import laplacian_lambda_max
                lambda_max = laplacian_lambda_max(graph)
            if isinstance(lambda_max, list):
                lambda_max = torch.Tensor(lambda_max).to(feat)
            if lambda_max.dim() == 1:
                lambda_max = lambda_max.unsqueeze(-1)
            from ...readout import broadcast_nodes
            lambda_max = broadcast_nodes(graph, lambda_max)
            re_norm = 2.0 / lambda_max
            X_0 = feat
            Xt = X_0
            if self._k > 1:
                h = unnLaplacian(X_0, D_invsqrt, graph)
                X_1 = -re_norm * h + X_0 * (re_norm - 1)
                Xt = torch.cat((Xt, X_1), 1)
            for _ in range(2, self._k):
                h = unnLaplacian(X_1, D_invsqrt, graph)
                X_i = -2 * re_norm * h + X_1 * 2 * (re_norm - 1) - X_0
                Xt = torch.cat((Xt, X_i), 1)
                X_1, X_0 = X_i, X_1
            h = self.linear(Xt)
            if self.activation:
                h = self.activation(h)
            return h


class GCN2Conv(nn.Module):
    r"""GCNII layer with initial residual and identity mapping (reference gcn2conv.py)."""

    def __init__(self, in_feats, layer, alpha=0.1, lambda_=1, project_initial_features=True,
                 allow_zero_in_degree=False, bias=True, activation=None):
        super().__init__()
        self._in_feats = in_feats
        self._project_initial_features = project_initial_features
        self.alpha = alpha
        self.beta = math.log(lambda_ / layer + 1)
        self._bias = bias
        self._activation = activation
        self._allow_zero_in_degree = allow_zero_in_degree
        self.weight1 = nn.Parameter(torch.Tensor(in_feats, in_feats))
        if self._project_initial_features:
            self.register_parameter("weight2", None)
        else:
            self.weight2 = nn.Parameter(torch.Tensor(in_feats, in_feats))
        if self._bias:
            self.bias = nn.Parameter(torch.Tensor(in_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.weight1)
        if not self._project_initial_features:
            nn.init.normal_(self.weight2)
        if self._bias:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, feat_0, edge_weight=None):
        with graph.local_scope():
            if not self._allow_zero_in_degree and bool((graph.in_degrees() == 0).any()):
                raise DGLError("There are 0-in-degree nodes in the graph; add self-loops or set allow_zero_in_degree.")
            if edge_weight is None:
                degs = graph.in_degrees().to(feat).clamp(min=1)
                norm = torch.pow(degs, -0.5).unsqueeze(-1)
                feat = feat * norm
                graph.ndata["h"] = feat
                graph.update_all(fn.copy_u("h", "m"), fn.sum("m", "h"))
                feat = graph.ndata.pop("h") * norm
            else:
                graph.ndata["h"] = feat
                graph.edata["_w"] = edge_weight
                graph.update_all(fn.u_mul_e("h", "_w", "m"), fn.sum("m", "h"))
                feat = graph.ndata.pop("h")
            if self._project_initial_features:
                rst = (1 - self.alpha) * feat + self.alpha * feat_0
                rst = torch.addmm(rst, rst, self.weight1, beta=(1 - self.beta), alpha=self.beta)
            else:
                rst = torch.addmm(feat, feat, self.weight1, beta=(1 - self.beta), alpha=self.beta * (1 - self.alpha))
                rst = rst + torch.addmm(feat_0, feat_0, self.weight2, beta=(1 - self.beta), alpha=self.beta * self.alpha)
            if self._bias:
                rst = rst + self.bias
            if self._activation is not None:
                rst = self._activation(rst)
            return rst


class GatedGraphConv(nn.Module):
    r"""Gated graph conv with GRU update and per-etype linear maps (reference gatedgraphconv.py)."""

    def __init__(self, in_feats, out_feats, n_steps, n_etypes, bias=True):
        super().__init__()
        self._in_feats, self._out_feats, self._n_steps, self._n_etypes = in_feats, out_feats, n_steps, n_etypes
        self.linears = nn.ModuleList([nn.Linear(out_feats, out_feats) for _ in range(n_etypes)])
        self.gru = nn.GRUCell(out_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        self.gru.reset_parameters()
        for linear in self.linears:
            nn.init.xavier_normal_(linear.weight, gain=gain)
            nn.init.zeros_(linear.bias)

    def set_allow_zero_in_degree(self, set_value):
        """Kept for API parity: the GRU update of this layer is well defined for nodes without in-edges (their
        aggregated message is zero), so nothing is checked either way (reference gatedgraphconv.py:98)."""
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, etypes=None):
        with graph.local_scope():
            assert graph.is_homogeneous, "not a homogeneous graph; convert it with to_homogeneous and pass in the edge type as argument"
            if self._n_etypes != 1:
                assert etypes.min() >= 0 and etypes.max() < self._n_etypes, "edge type indices out of range"
            zero_pad = feat.new_zeros((feat.shape[0], self._out_feats - feat.shape[1]))
            feat = torch.cat([feat, zero_pad], -1)
            for _ in range(self._n_steps):
                if self._n_etypes == 1 and etypes is None:
                    graph.ndata["h"] = self.linears[0](feat)
                    graph.update_all(fn.copy_u("h", "m"), fn.sum("m", "a"))
                    a = graph.ndata.pop("a")
                else:
                    graph.ndata["h"] = feat
                    a = 0
                    for i in range(self._n_etypes):
                        eids = torch.nonzero(etypes == i, as_tuple=False).view(-1).type(graph.idtype)
                        if len(eids) > 0:
                            sub = graph.edge_subgraph(eids, relabel_nodes=False)
                            sub.ndata["hh"] = self.linears[i](feat)
                            sub.update_all(fn.copy_u("hh", "m"), fn.sum("m", "a"))
                            a = a + sub.ndata["a"]
                feat = self.gru(a, feat)
            return feat


class GatedGCNConv(nn.Module):
    r"""Gated GCN (Bresson & Laurent) with edge features (reference gatedgcnconv.py)."""

    def __init__(self, input_feats, edge_feats, output_feats, dropout=0, batch_norm=True, residual=True,
                 activation=F.relu):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.batch_norm = batch_norm
        self.residual = residual
        if input_feats != output_feats or edge_feats != output_feats:
            self.residual = False
        self.A = nn.Linear(input_feats, output_feats, bias=True)
        self.B = nn.Linear(input_feats, output_feats, bias=True)
        self.C = nn.Linear(edge_feats, output_feats, bias=True)
        self.D = nn.Linear(input_feats, output_feats, bias=True)
        self.E = nn.Linear(input_feats, output_feats, bias=True)
        self.bn_node = nn.BatchNorm1d(output_feats)
        self.bn_edge = nn.BatchNorm1d(output_feats)
        self.activation = activation

    def forward(self, graph, feat, edge_feat):
        with graph.local_scope():
            h_in, e_in = feat, edge_feat
            graph.ndata["Ah"] = self.A(feat)
            graph.ndata["Bh"] = self.B(feat)
            graph.ndata["Dh"] = self.D(feat)
            graph.ndata["Eh"] = self.E(feat)
            graph.edata["Ce"] = self.C(edge_feat)
            graph.apply_edges(fn.u_add_v("Dh", "Eh", "DEh"))
            graph.edata["e"] = graph.edata["DEh"] + graph.edata["Ce"]
            graph.edata["sigma"] = torch.sigmoid(graph.edata["e"])
            graph.update_all(fn.u_mul_e("Bh", "sigma", "m"), fn.sum("m", "sum_sigma_h"))
            graph.update_all(fn.copy_e("sigma", "m"), fn.sum("m", "sum_sigma"))
            graph.ndata["h"] = graph.ndata["Ah"] + graph.ndata["sum_sigma_h"] / (graph.ndata["sum_sigma"] + 1e-6)
            h, e = graph.ndata["h"], graph.edata["e"]
            if self.batch_norm:
                h, e = self.bn_node(h), self.bn_edge(e)
            if self.activation:
                h, e = self.activation(h), self.activation(e)
            if self.residual:
                h, e = h_in + h, e_in + e
            return self.dropout(h), self.dropout(e)


class ShiftedSoftplus(nn.Module):
    def __init__(self, beta=1, shift=2, threshold=20):
        super().__init__()
        self.shift = shift
        self.softplus = nn.Softplus(beta=beta, threshold=threshold)

    def forward(self, inputs):
        return self.softplus(inputs) - math.log(float(self.shift))


class CFConv(nn.Module):
    r"""SchNet continuous-filter convolution (reference cfconv.py)."""

    def __init__(self, node_in_feats, edge_in_feats, hidden_feats, out_feats):
        super().__init__()
        self.project_edge = nn.Sequential(nn.Linear(edge_in_feats, hidden_feats), ShiftedSoftplus(),
                                          nn.Linear(hidden_feats, hidden_feats), ShiftedSoftplus())
        self.project_node = nn.Linear(node_in_feats, hidden_feats)
        self.project_out = nn.Sequential(nn.Linear(hidden_feats, out_feats), ShiftedSoftplus())

    def forward(self, g, node_feats, edge_feats):
        with g.local_scope():
            if isinstance(node_feats, tuple):
                node_feats_src, _ = node_feats
            else:
                node_feats_src = node_feats
            g.srcdata["hv"] = self.project_node(node_feats_src)
            g.edata["he"] = self.project_edge(edge_feats)
            g.update_all(fn.u_mul_e("hv", "he", "m"), fn.sum("m", "h"))
            return self.project_out(g.dstdata["h"])


class NNConv(nn.Module):
    r"""Edge-conditioned convolution (MPNN): W_e = f(e) (reference nnconv.py)."""

    def __init__(self, in_feats, out_feats, edge_func, aggregator_type="mean", residual=False, bias=True):
        super().__init__()
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats = out_feats
        self.edge_func = edge_func
        if aggregator_type not in ("sum", "mean", "max"):
            raise KeyError("Aggregator type {} not recognized: ".format(aggregator_type))
        self.reducer = getattr(fn, aggregator_type)
        self._aggre_type = aggregator_type
        if residual:
            if self._in_dst_feats != out_feats:
                self.res_fc = nn.Linear(self._in_dst_feats, out_feats, bias=False)
            else:
                self.res_fc = nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        if self.bias is not None:
            nn.init.zeros_(self.bias)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def forward(self, graph, feat, efeat):
        with graph.local_scope():
            feat_src, feat_dst = expand_as_pair(feat, graph)
            graph.srcdata["h"] = feat_src.unsqueeze(-1)
            graph.edata["w"] = self.edge_func(efeat).view(-1, self._in_src_feats, self._out_feats)
            graph.update_all(fn.u_mul_e("h", "w", "m"), self.reducer("m", "neigh"))
            rst = graph.dstdata["neigh"].sum(dim=1)
            if self.res_fc is not None:
                rst = rst + self.res_fc(feat_dst)
            if self.bias is not None:
                rst = rst + self.bias
            return rst


class GMMConv(nn.Module):
    r"""Gaussian mixture model convolution (MoNet) (reference gmmconv.py)."""

    def __init__(self, in_feats, out_feats, dim, n_kernels, aggregator_type="sum", residual=False, bias=True,
                 allow_zero_in_degree=False):
        super().__init__()
        self._in_src_feats, self._in_dst_feats = expand_as_pair(in_feats)
        self._out_feats, self._dim, self._n_kernels = out_feats, dim, n_kernels
        self._allow_zero_in_degree = allow_zero_in_degree
        if aggregator_type not in ("sum", "mean", "max"):
            raise KeyError("Aggregator type {} not recognized.".format(aggregator_type))
        self._reducer = getattr(fn, aggregator_type)
        self.mu = nn.Parameter(torch.Tensor(n_kernels, dim))
        self.inv_sigma = nn.Parameter(torch.Tensor(n_kernels, dim))
        self.fc = nn.Linear(self._in_src_feats, n_kernels * out_feats, bias=False)
        if residual:
            if self._in_dst_feats != out_feats:
                self.res_fc = nn.Linear(self._in_dst_feats, out_feats, bias=False)
            else:
                self.res_fc = nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)
        nn.init.normal_(self.mu.data, 0, 0.1)
        nn.init.constant_(self.inv_sigma.data, 1)
        if self.bias is not None:
            nn.init.zeros_(self.bias.data)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, pseudo):
        with graph.local_scope():
            if not self._allow_zero_in_degree and bool((graph.in_degrees() == 0).any()):
                raise DGLError("There are 0-in-degree nodes in the graph; add self-loops or set allow_zero_in_degree.")
            feat_src, feat_dst = expand_as_pair(feat, graph)
            graph.srcdata["h"] = self.fc(feat_src).view(-1, self._n_kernels, self._out_feats)
            E = graph.num_edges()
            gaussian = -0.5 * ((pseudo.view(E, 1, self._dim) - self.mu.view(1, self._n_kernels, self._dim)) ** 2)
            gaussian = gaussian * (self.inv_sigma.view(1, self._n_kernels, self._dim) ** 2)
            gaussian = torch.exp(gaussian.sum(dim=-1, keepdim=True))
            graph.edata["w"] = gaussian
            graph.update_all(fn.u_mul_e("h", "w", "m"), self._reducer("m", "h"))
            rst = graph.dstdata["h"].sum(1)
            if self.res_fc is not None:
                rst = rst + self.res_fc(feat_dst)
            if self.bias is not None:
                rst = rst + self.bias
            return rst


class EdgeConv(nn.Module):
    r"""DGCNN EdgeConv: h_i' = max_j (Theta (h_j - h_i) + Phi h_i) (reference edgeconv.py)."""

    def __init__(self, in_feat, out_feat, batch_norm=False, allow_zero_in_degree=False):
        super().__init__()
        self.batch_norm = batch_norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.theta = nn.Linear(in_feat, out_feat)
        self.phi = nn.Linear(in_feat, out_feat)
        if batch_norm:
            self.bn = nn.BatchNorm1d(out_feat)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, g, feat):
        with g.local_scope():
            if not self._allow_zero_in_degree and bool((g.in_degrees() == 0).any()):
                raise DGLError("There are 0-in-degree nodes in the graph; add self-loops or set allow_zero_in_degree.")
            h_src, h_dst = expand_as_pair(feat, g)
            g.srcdata["x"] = h_src
            g.dstdata["x"] = h_dst
            g.apply_edges(fn.v_sub_u("x", "x", "theta"))
            g.edata["theta"] = self.theta(g.edata["theta"])
            g.dstdata["phi"] = self.phi(g.dstdata["x"])
            if not self.batch_norm:
                g.update_all(fn.e_add_v("theta", "phi", "e"), fn.max("e", "x"))
            else:
                g.apply_edges(fn.e_add_v("theta", "phi", "e"))
                g.edata["e"] = self.bn(g.edata["e"])
                g.update_all(fn.copy_e("e", "e"), fn.max("e", "x"))
            return g.dstdata["x"]


class EGNNConv(nn.Module):
    r"""E(n)-equivariant GNN layer (reference egnnconv.py)."""

    def __init__(self, in_size, hidden_size, out_size, edge_feat_size=0):
        super().__init__()
        self.in_size, self.hidden_size, self.out_size, self.edge_feat_size = in_size, hidden_size, out_size, edge_feat_size
        act_fn = nn.SiLU()
        self.edge_mlp = nn.Sequential(nn.Linear(in_size * 2 + edge_feat_size + 1, hidden_size), act_fn,
                                      nn.Linear(hidden_size, hidden_size), act_fn)
        self.node_mlp = nn.Sequential(nn.Linear(in_size + hidden_size, hidden_size), act_fn,
                                      nn.Linear(hidden_size, out_size))
        self.coord_mlp = nn.Sequential(nn.Linear(hidden_size, hidden_size), act_fn, nn.Linear(hidden_size, 1, bias=False))

    def message(self, edges):
        if self.edge_feat_size > 0:
            f = torch.cat([edges.src["h"], edges.dst["h"], edges.data["radial"], edges.data["a"]], dim=-1)
        else:
            f = torch.cat([edges.src["h"], edges.dst["h"], edges.data["radial"]], dim=-1)
        msg_h = self.edge_mlp(f)
        msg_x = self.coord_mlp(msg_h) * edges.data["x_diff"]
        return {"msg_x": msg_x, "msg_h": msg_h}

    def forward(self, graph, node_feat, coord_feat, edge_feat=None):
        with graph.local_scope():
            graph.ndata["h"] = node_feat
            graph.ndata["x"] = coord_feat
            if self.edge_feat_size > 0:
                assert edge_feat is not None, "Edge features must be provided."
                graph.edata["a"] = edge_feat
            graph.apply_edges(fn.u_sub_v("x", "x", "x_diff"))
            graph.edata["radial"] = graph.edata["x_diff"].square().sum(dim=1).unsqueeze(-1)
            graph.edata["x_diff"] = graph.edata["x_diff"] / (graph.edata["radial"].sqrt() + 1e-30)
            graph.apply_edges(self.message)
            graph.update_all(fn.copy_e("msg_x", "m"), fn.mean("m", "x_neigh"))
            graph.update_all(fn.copy_e("msg_h", "m"), fn.sum("m", "h_neigh"))
            h_neigh, x_neigh = graph.ndata["h_neigh"], graph.ndata["x_neigh"]
            h = self.node_mlp(torch.cat([node_feat, h_neigh], dim=-1))
            x = coord_feat + x_neigh
            return h, x


class TWIRLSConv(nn.Module):
    r"""Together-with-IRLS unfolding layer (reference twirlsconv.py), simplified to the propagation +
    optional attention reweighting steps: Y <- (1-alpha*...) iterations of normalised propagation."""

    def __init__(self, input_d, output_d, hidden_d, prop_step, num_mlp_before=1, num_mlp_after=1, norm="none",
                 precond=True, alp=0, lam=1, attention=False, tau=0.2, T=-1, p=1, use_eta=False, attn_bef=False,
                 dropout=0.0, attn_dropout=0.0, inp_dropout=0.0):
        super().__init__()
        self.input_d, self.output_d, self.hidden_d = input_d, output_d, hidden_d
        self.prop_step, self.lam, self.precond = prop_step, lam, precond
        self.alp = alp if alp > 0 else 1.0 / (lam + 1)
        self.attention, self.tau, self.T, self.p = attention, tau, T, p
        self.attn_step = prop_step // 2 if attention else -1

        def mlp(i, o, n):
            if n == 0:
                return nn.Identity()
            layers = []
            d = i
            for k in range(n):
                od = o if k == n - 1 else hidden_d
                layers.append(nn.Linear(d, od))
                if k != n - 1:
                    layers += [nn.ReLU(), nn.Dropout(dropout)]
                d = od
            return nn.Sequential(*layers)

        self.size_bef = hidden_d if num_mlp_after > 0 else output_d
        self.mlp_bef = mlp(input_d, self.size_bef, num_mlp_before)
        self.mlp_aft = mlp(self.size_bef if num_mlp_before > 0 else input_d, output_d, num_mlp_after)
        self.inp_dropout = nn.Dropout(inp_dropout)
        self.attn_dropout = nn.Dropout(attn_dropout)

    def forward(self, graph, feat):
        with graph.local_scope():
            x = self.mlp_bef(self.inp_dropout(feat))
            deg = graph.in_degrees().to(x).clamp(min=1).unsqueeze(-1)
            Y = x
            w = torch.ones(graph.num_edges(), 1, device=x.device, dtype=x.dtype)
            for k in range(self.prop_step):
                if k == self.attn_step:
                    graph.ndata["_y"] = Y
                    graph.apply_edges(fn.u_sub_v("_y", "_y", "_d"))
                    d2 = graph.edata["_d"].square().sum(-1, keepdim=True)
                    w = (d2.clamp(min=self.tau ** 2)) ** (self.p / 2 - 1)
                    if self.T > 0:
                        w = torch.where(d2 > self.T ** 2, torch.zeros_like(w), w)
                    w = self.attn_dropout(w)
                graph.ndata["_y"] = Y / torch.sqrt(deg) if not self.precond else Y
                graph.edata["_w"] = w
                graph.update_all(fn.u_mul_e("_y", "_w", "m"), fn.sum("m", "_ay"))
                AY = graph.ndata["_ay"]
                if self.precond:
                    Y = (1 - self.alp) * Y + self.alp * self.lam * AY / (self.lam * deg + 1) + self.alp * x / (self.lam * deg + 1)
                else:
                    Y = (1 - self.alp) * Y + self.alp * self.lam * AY / torch.sqrt(deg) + self.alp * x
            return self.mlp_aft(Y)
