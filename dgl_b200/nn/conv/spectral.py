"""Propagation-filter, isomorphism and edge-conditioned convolutions.

GINConv / GINEConv, SGConv / TAGConv / APPNPConv / ChebConv / GCN2Conv (polynomial filters of the normalised adjacency),
GatedGraphConv / GatedGCNConv, CFConv / NNConv / GMMConv / EdgeConv / EGNNConv, TWIRLSConv.

Constructor signatures and parameter names follow /root/reference/python/dgl/nn/pytorch/conv/<layer>.py (checkpoints
interchange; ``tests/test_nn_reference_parity.py`` checks values and gradients against the unmodified reference). The
forward passes are written on the tensor-level operators of ``nn/_mp.py``:

* every degree normalisation is folded into cached per-edge coefficients, so one propagation step is ONE weighted
  g-SpMM (``_mp.sym_norm_propagate``) instead of scale -> SpMM -> scale;
* linear maps that the reference applies per EDGE are hoisted to the nodes whenever the message is linear in an
  endpoint (EdgeConv: ``Theta (x_v - x_u)`` = ``(Theta x)_v - (Theta x)_u``, EGNN: the first edge-MLP layer splits into
  per-endpoint GEMMs), which turns E x d GEMMs into N x d ones and removes [E, in] intermediates;
* typed per-edge transforms (GatedGraphConv) run as one grouped GEMM (``ops.gather_mm``) instead of a Python loop of
  ``send_and_recv`` over edge-type subsets.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...base import DGLError
from .. import _mp
from .graphconv import EdgeWeightNorm
from .parts import MLP, TWIRLSUnfoldingAndAttention

_REDUCERS = ("sum", "max", "mean")


def _eps_holder(module, init_eps, learn_eps):
    value = torch.tensor([float(init_eps)])
    if learn_eps:
        module.eps = nn.Parameter(value)
    else:
        module.register_buffer("eps", value)


def _both_norm_or_weights(g, dtype, edge_weight, mode="in_both"):
    """Edge coefficients of one propagation step: the structural symmetric normalisation (``in_both``: the in-degree
    vector on both ends, as these filters are defined on undirected graphs; ``both``: out-degree of the source), or the
    caller's weights symmetric-normalised by their weighted degrees."""
    if edge_weight is None:
        return _mp.norm_coefficients(g, mode, dtype)
    return EdgeWeightNorm("both")(g, edge_weight)


# =====================================================================================================================
#  isomorphism networks
# =====================================================================================================================
class GINConv(nn.Module):
    r"""``h_i' = f( (1 + eps) h_i + aggregate_j w_ji h_j )`` with aggregate in sum | max | mean."""

    def __init__(self, apply_func=None, aggregator_type="sum", init_eps=0, learn_eps=False, activation=None):
        super().__init__()
        if aggregator_type not in _REDUCERS:
            raise KeyError(f"Aggregator type {aggregator_type} not recognized.")
        self.apply_func, self._aggregator_type, self.activation = apply_func, aggregator_type, activation
        _eps_holder(self, init_eps, learn_eps)

    def forward(self, graph, feat, edge_weight=None):
        if _mp.is_capacity_block(graph):                              # whole-step CUDA graph: ELL kernels, no size reads
            x = feat[0] if isinstance(feat, (tuple, list)) else feat
            x_dst, pooled = _mp.capacity_aggregate(graph, x, self._aggregator_type)
            out = torch.addcmul(pooled, x_dst, (1 + self.eps).to(x_dst.dtype))
            if self.apply_func is not None:
                out = self.apply_func(out)
            return out if self.activation is None else self.activation(out)
        x_src, x_dst = _mp.src_dst(graph, feat)
        if edge_weight is not None and edge_weight.shape[0] != graph.num_edges():
            raise DGLError("edge_weight must have one entry per edge")
        pooled = _mp.gather(graph, x_src, self._aggregator_type, edge_weight)
        out = torch.addcmul(pooled, x_dst, (1 + self.eps).to(x_dst.dtype))
        if self.apply_func is not None:
            out = self.apply_func(out)
        return out if self.activation is None else self.activation(out)


class GINEConv(nn.Module):
    r"""``h_i' = f( (1 + eps) h_i + sum_j relu(h_j + e_ji) )``."""

    def __init__(self, apply_func=None, init_eps=0, learn_eps=False):
        super().__init__()
        self.apply_func = apply_func
        _eps_holder(self, init_eps, learn_eps)

    def forward(self, graph, node_feat, edge_feat):
        x_src, x_dst = _mp.src_dst(graph, node_feat)
        msg = F.relu(ops.gsddmm(graph, "add", x_src, edge_feat, "u", "e"))
        out = torch.addcmul(_mp.edge_reduce(graph, msg, "sum"), x_dst, (1 + self.eps).to(x_dst.dtype))
        return out if self.apply_func is None else self.apply_func(out)


# =====================================================================================================================
#  polynomial filters of the normalised adjacency
# =====================================================================================================================
class SGConv(nn.Module):
    r"""Simplified GCN: ``H = (D^-1/2 A D^-1/2)^k X Theta`` (optionally cached after the first call)."""

    def __init__(self, in_feats, out_feats, k=1, cached=False, bias=True, norm=None, allow_zero_in_degree=False):
        super().__init__()
        self.fc = nn.Linear(in_feats, out_feats, bias=bias)
        self._cached, self._cached_h, self._k, self.norm = cached, None, k, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.fc.weight)
        if self.fc.bias is not None:
            nn.init.zeros_(self.fc.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, edge_weight=None):
        _mp.require_in_edges(graph, self._allow_zero_in_degree, "SGConv")
        if self._cached_h is not None:
            h = self._cached_h
        else:
            coeff = _both_norm_or_weights(graph, feat.dtype, edge_weight)
            h = feat
            for _ in range(self._k):
                h = _mp.gather_sum(graph, h, coeff)
            if self.norm is not None:
                h = self.norm(h)
            if self._cached:
                self._cached_h = h
        return self.fc(h)


class TAGConv(nn.Module):
    r"""Topology-adaptive GCN: ``H = [X | ÂX | ... | Â^k X] Theta``."""

    def __init__(self, in_feats, out_feats, k=2, bias=True, activation=None):
        super().__init__()
        self._in_feats, self._out_feats, self._k, self._activation = in_feats, out_feats, k, activation
        self.lin = nn.Linear(in_feats * (k + 1), out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_normal_(self.lin.weight, gain=nn.init.calculate_gain("relu"))

    def forward(self, graph, feat, edge_weight=None):
        if graph.is_block or not graph.is_homogeneous:
            raise DGLError("Graph is not homogeneous")
        coeff = _both_norm_or_weights(graph, feat.dtype, edge_weight)
        hops = [feat]
        for _ in range(self._k):
            hops.append(_mp.gather_sum(graph, hops[-1], coeff))
        out = self.lin(torch.cat(hops, dim=-1))
        return out if self._activation is None else self._activation(out)


class APPNPConv(nn.Module):
    r"""Approximate personalised PageRank propagation ``H_{t+1} = (1 - alpha) Â H_t + alpha H_0`` for k steps."""

    def __init__(self, k, alpha, edge_drop=0.0):
        super().__init__()
        self._k, self._alpha = k, alpha
        self.edge_drop = nn.Dropout(edge_drop)

    def forward(self, graph, feat, edge_weight=None):
        coeff = _both_norm_or_weights(graph, feat.dtype, edge_weight, mode="both")
        h, teleport = feat, self._alpha * feat
        for _ in range(self._k):
            w = self.edge_drop(coeff) if self.training and self.edge_drop.p > 0 else coeff
            h = (1 - self._alpha) * _mp.gather_sum(graph, h, w) + teleport
        return h


class ChebConv(nn.Module):
    r"""Chebyshev spectral convolution of order k on the rescaled Laplacian ``L~ = 2 L / lambda_max - I``."""

    def __init__(self, in_feats, out_feats, k, activation=F.relu, bias=True):
        super().__init__()
        self._k, self._in_feats, self._out_feats, self.activation = k, in_feats, out_feats, activation
        self.linear = nn.Linear(k * in_feats, out_feats, bias)

    def _node_scale(self, graph, lambda_max, like):
        if lambda_max is None:
            from ...transforms import laplacian_lambda_max
            lambda_max = laplacian_lambda_max(graph)
        if not torch.is_tensor(lambda_max):
            lambda_max = torch.as_tensor(lambda_max, dtype=like.dtype)
        lambda_max = lambda_max.to(like)
        if lambda_max.dim() == 1:
            lambda_max = lambda_max.unsqueeze(-1)
        from ...readout import broadcast_nodes
        return 2.0 / broadcast_nodes(graph, lambda_max)          # one rescaling factor per node of a batched graph

    def forward(self, graph, feat, lambda_max=None):
        s = self._node_scale(graph, lambda_max, feat)
        coeff = _mp.norm_coefficients(graph, "in_both", feat.dtype)

        def rescaled_laplacian(x):                                # L~ x = s (x - Â x) - x
            return (s - 1) * x - s * _mp.gather_sum(graph, x, coeff)

        terms = [feat]
        if self._k > 1:
            terms.append(rescaled_laplacian(feat))
        for _ in range(2, self._k):                               # T_i = 2 L~ T_{i-1} - T_{i-2}
            terms.append(2 * rescaled_laplacian(terms[-1]) - terms[-2])
        out = self.linear(torch.cat(terms, dim=1))
        return out if self.activation is None else self.activation(out)


class GCN2Conv(nn.Module):
    r"""GCNII layer: initial residual (alpha) + identity mapping (beta = log(lambda / layer + 1))."""

    def __init__(self, in_feats, layer, alpha=0.1, lambda_=1, project_initial_features=True,
                 allow_zero_in_degree=False, bias=True, activation=None):
        super().__init__()
        self._in_feats, self._project_initial_features = in_feats, project_initial_features
        self.alpha, self.beta = alpha, math.log(lambda_ / layer + 1)
        self._bias, self._activation, self._allow_zero_in_degree = bias, activation, allow_zero_in_degree
        self.weight1 = nn.Parameter(torch.empty(in_feats, in_feats))
        self.weight2 = None if project_initial_features else nn.Parameter(torch.empty(in_feats, in_feats))
        self.bias = nn.Parameter(torch.empty(in_feats)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.normal_(self.weight1)
        if self.weight2 is not None:
            nn.init.normal_(self.weight2)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def _identity_map(self, x, w):
        return torch.addmm(x, x, w, beta=1 - self.beta, alpha=self.beta)   # (1 - beta) x + beta x W

    def forward(self, graph, feat, feat_0, edge_weight=None):
        _mp.require_in_edges(graph, self._allow_zero_in_degree, "GCN2Conv")
        coeff = _both_norm_or_weights(graph, feat.dtype, edge_weight)
        smooth = (1 - self.alpha) * _mp.gather_sum(graph, feat, coeff)
        init = self.alpha * feat_0[: smooth.shape[0]]
        if self._project_initial_features:
            out = self._identity_map(smooth + init, self.weight1)
        else:
            out = self._identity_map(smooth, self.weight1) + self._identity_map(init, self.weight2)
        if self.bias is not None:
            out = out + self.bias
        return out if self._activation is None else self._activation(out)

    def extra_repr(self):
        s = f"in={self._in_feats}, alpha={self.alpha}, beta={self.beta}"
        if self._activation is not None:
            s += f", activation={self._activation}"
        return s


# =====================================================================================================================
#  gated recurrences
# =====================================================================================================================
class GatedGraphConv(nn.Module):
    r"""GGNN: ``a_i = sum_j W_{e_ji} h_j``, ``h_i <- GRU(a_i, h_i)`` repeated ``n_steps`` times."""

    def __init__(self, in_feats, out_feats, n_steps, n_etypes, bias=True):
        super().__init__()
        self._in_feats, self._out_feats, self._n_steps, self._n_etypes = in_feats, out_feats, n_steps, n_etypes
        self.linears = nn.ModuleList([nn.Linear(out_feats, out_feats) for _ in range(n_etypes)])
        self.gru = nn.GRUCell(out_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        self.gru.reset_parameters()
        for lin in self.linears:
            nn.init.xavier_normal_(lin.weight, gain=gain)
            nn.init.zeros_(lin.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def _incoming(self, graph, h, etypes):
        if self._n_etypes == 1 and etypes is None:
            return _mp.gather_sum(graph, self.linears[0](h))
        # one grouped GEMM for all edge types: message_e = W_{t(e)} h_src(e) + b_{t(e)}
        W = torch.stack([lin.weight.t() for lin in self.linears])              # [T, in, out]
        b = torch.stack([lin.bias for lin in self.linears])                    # [T, out]
        t = etypes.long()
        msg = ops.gather_mm(ops.copy_u(graph, h), W, idx_b=t) + b[t]
        return _mp.edge_reduce(graph, msg, "sum")

    def forward(self, graph, feat, etypes=None):
        if not graph.is_homogeneous:
            raise DGLError("not a homograph; convert it with to_homogeneous and pass in the edge type as argument")
        h = F.pad(feat, (0, self._out_feats - feat.shape[1]))
        for _ in range(self._n_steps):
            h = self.gru(self._incoming(graph, h, etypes), h)
        return h


class GatedGCNConv(nn.Module):
    r"""Residual gated GCN (Bresson & Laurent): edge gates ``sigma(D h_u + E h_v + C e)`` weight the ``B h_u`` messages."""

    def __init__(self, input_feats, edge_feats, output_feats, dropout=0, batch_norm=True, residual=True,
                 activation=F.relu):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.batch_norm, self.activation = batch_norm, activation
        self.residual = residual and input_feats == output_feats
        for name, width in (("A", input_feats), ("B", input_feats), ("C", edge_feats), ("D", input_feats), ("E", input_feats)):
            setattr(self, name, nn.Linear(width, output_feats, bias=True))
        self.bn_node = nn.BatchNorm1d(output_feats)
        self.bn_edge = nn.BatchNorm1d(output_feats)

    def forward(self, graph, feat, edge_feat):
        h_in, e_in = feat, edge_feat
        gate_logit = _mp.pair_op(graph, "add", self.D(feat), self.E(feat)) + self.C(edge_feat)
        gate = torch.sigmoid(gate_logit)
        num = _mp.gather_sum(graph, self.B(feat), gate)                         # sum_e gate * (B h)_u
        den = _mp.edge_reduce(graph, gate, "sum")
        h, e = self.A(feat) + num / (den + 1e-6), gate_logit
        if self.batch_norm:
            h, e = self.bn_node(h), self.bn_edge(e)
        if self.activation is not None:
            h, e = self.activation(h), self.activation(e)
        if self.residual:
            h, e = h_in + h, e_in + e
        return self.dropout(h), self.dropout(e)


# =====================================================================================================================
#  edge-conditioned filters
# =====================================================================================================================
class ShiftedSoftplus(nn.Module):
    r"""``softplus(x) - log(shift)``."""

    def __init__(self, beta=1, shift=2, threshold=20):
        super().__init__()
        self.shift = shift
        self.softplus = nn.Softplus(beta=beta, threshold=threshold)

    def forward(self, inputs):
        return self.softplus(inputs) - math.log(float(self.shift))


class CFConv(nn.Module):
    r"""SchNet continuous-filter convolution: ``h_i' = f_out( sum_j f_node(h_j) * f_edge(e_ji) )``."""

    def __init__(self, node_in_feats, edge_in_feats, hidden_feats, out_feats):
        super().__init__()
        self.project_edge = nn.Sequential(nn.Linear(edge_in_feats, hidden_feats), ShiftedSoftplus(),
                                          nn.Linear(hidden_feats, hidden_feats), ShiftedSoftplus())
        self.project_node = nn.Linear(node_in_feats, hidden_feats)
        self.project_out = nn.Sequential(nn.Linear(hidden_feats, out_feats), ShiftedSoftplus())

    def forward(self, g, node_feats, edge_feats):
        x_src, _ = _mp.src_dst(g, node_feats)
        return self.project_out(_mp.gather_sum(g, self.project_node(x_src), self.project_edge(edge_feats)))


def _residual_and_bias(module, in_dst, out_feats, residual, bias):
    if not residual:
        module.register_buffer("res_fc", None)
    elif in_dst != out_feats:
        module.res_fc = nn.Linear(in_dst, out_feats, bias=False)
    else:
        module.res_fc = nn.Identity()
    if bias:
        module.bias = nn.Parameter(torch.zeros(out_feats))
    else:
        module.register_buffer("bias", None)


class NNConv(nn.Module):
    r"""Edge-conditioned convolution (MPNN): every edge owns a matrix ``Theta(e_ji)`` produced by ``edge_func``."""

    def __init__(self, in_feats, out_feats, edge_func, aggregator_type="mean", residual=False, bias=True):
        super().__init__()
        self._in_src_feats, self._in_dst_feats = (in_feats, in_feats) if isinstance(in_feats, int) else in_feats
        self._out_feats, self.edge_func = out_feats, edge_func
        if aggregator_type not in _REDUCERS:
            raise KeyError(f"Aggregator type {aggregator_type} not recognized: ")
        self._aggre_type = aggregator_type
        _residual_and_bias(self, self._in_dst_feats, out_feats, residual, bias)
        self.reset_parameters()

    def reset_parameters(self):
        if self.bias is not None:
            nn.init.zeros_(self.bias)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=nn.init.calculate_gain("relu"))

    def forward(self, graph, feat, efeat):
        x_src, x_dst = _mp.src_dst(graph, feat)
        theta = self.edge_func(efeat).view(-1, self._in_src_feats, self._out_feats)
        msg = torch.bmm(ops.copy_u(graph, x_src).unsqueeze(1), theta).squeeze(1)   # per-edge vector-matrix product
        out = _mp.edge_reduce(graph, msg, self._aggre_type)
        if self.res_fc is not None:
            out = out + self.res_fc(x_dst)
        return out if self.bias is None else out + self.bias


class GMMConv(nn.Module):
    r"""MoNet: Gaussian mixture kernels on pseudo-coordinates weight per-kernel projections of the neighbours."""

    def __init__(self, in_feats, out_feats, dim, n_kernels, aggregator_type="sum", residual=False, bias=True,
                 allow_zero_in_degree=False):
        super().__init__()
        self._in_src_feats, self._in_dst_feats = (in_feats, in_feats) if isinstance(in_feats, int) else in_feats
        self._out_feats, self._dim, self._n_kernels = out_feats, dim, n_kernels
        self._allow_zero_in_degree = allow_zero_in_degree
        if aggregator_type not in _REDUCERS:
            raise KeyError(f"Aggregator type {aggregator_type} not recognized.")
        self._reducer = aggregator_type
        self.mu = nn.Parameter(torch.empty(n_kernels, dim))
        self.inv_sigma = nn.Parameter(torch.empty(n_kernels, dim))
        self.fc = nn.Linear(self._in_src_feats, n_kernels * out_feats, bias=False)
        _residual_and_bias(self, self._in_dst_feats, out_feats, residual, bias)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)
        nn.init.normal_(self.mu, 0, 0.1)
        nn.init.constant_(self.inv_sigma, 1)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, pseudo):
        _mp.require_in_edges(graph, self._allow_zero_in_degree, "GMMConv")
        x_src, x_dst = _mp.src_dst(graph, feat)
        proj = self.fc(x_src).view(-1, self._n_kernels, self._out_feats)
        z = (pseudo.unsqueeze(1) - self.mu.unsqueeze(0)) * self.inv_sigma.unsqueeze(0)       # [E, K, dim]
        kernel_w = torch.exp(-0.5 * z.square().sum(-1))                                       # [E, K]
        out = _mp.gather(graph, proj, self._reducer, kernel_w.unsqueeze(-1)).sum(1)
        if self.res_fc is not None:
            out = out + self.res_fc(x_dst)
        return out if self.bias is None else out + self.bias


class EdgeConv(nn.Module):
    r"""DGCNN edge convolution ``h_i' = max_j [ Theta (h_i - h_j) + Phi h_i ]`` over the in-edges j -> i.

    ``Theta`` is applied to the NODES: without batch norm the layer needs no per-edge tensor at all,
    ``max_j [(Theta h)_i - (Theta h)_j] = (Theta h)_i - min_j (Theta h)_j``."""

    def __init__(self, in_feat, out_feat, batch_norm=False, allow_zero_in_degree=False):
        super().__init__()
        self.batch_norm, self._allow_zero_in_degree = batch_norm, allow_zero_in_degree
        self.theta = nn.Linear(in_feat, out_feat)
        self.phi = nn.Linear(in_feat, out_feat)
        if batch_norm:
            self.bn = nn.BatchNorm1d(out_feat)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, g, feat):
        _mp.require_in_edges(g, self._allow_zero_in_degree, "EdgeConv")
        x_src, x_dst = _mp.src_dst(g, feat)
        t_src = F.linear(x_src, self.theta.weight)                       # (Theta h)_j, bias cancels in the difference
        centre = F.linear(x_dst, self.theta.weight) + self.theta.bias + self.phi(x_dst)
        if not self.batch_norm:
            has_nbr = (g.in_degrees() > 0).to(centre.dtype).unsqueeze(-1)  # the max over an empty set is 0
            return (centre - _mp.gather_min(g, t_src)) * has_nbr
        e = ops.gsddmm(g, "sub", centre, t_src, "v", "u")
        return _mp.edge_reduce(g, self.bn(e), "max")


class EGNNConv(nn.Module):
    r"""E(n)-equivariant layer: messages from (h_i, h_j, ||x_i - x_j||^2, e_ij); coordinates move along x_i - x_j."""

    def __init__(self, in_size, hidden_size, out_size, edge_feat_size=0):
        super().__init__()
        self.in_size, self.hidden_size, self.out_size, self.edge_feat_size = in_size, hidden_size, out_size, edge_feat_size
        act_fn = nn.SiLU()
        self.edge_mlp = nn.Sequential(nn.Linear(in_size * 2 + edge_feat_size + 1, hidden_size), act_fn,
                                      nn.Linear(hidden_size, hidden_size), act_fn)
        self.node_mlp = nn.Sequential(nn.Linear(in_size + hidden_size, hidden_size), act_fn,
                                      nn.Linear(hidden_size, out_size))
        self.coord_mlp = nn.Sequential(nn.Linear(hidden_size, hidden_size), act_fn, nn.Linear(hidden_size, 1, bias=False))

    def _first_edge_layer(self, graph, h, radial, edge_feat):
        """edge_mlp[0]([h_u | h_v | radial | e]) with the endpoint blocks of the weight applied per NODE."""
        lin, d = self.edge_mlp[0], self.in_size
        W = lin.weight
        z = _mp.pair_op(graph, "add", F.linear(h, W[:, :d]), F.linear(h, W[:, d:2 * d]))
        z = z + radial * W[:, 2 * d].unsqueeze(0) + lin.bias
        if self.edge_feat_size > 0:
            z = z + F.linear(edge_feat, W[:, 2 * d + 1:])
        return z

    def forward(self, graph, node_feat, coord_feat, edge_feat=None):
        if self.edge_feat_size > 0 and edge_feat is None:
            raise AssertionError("Edge features must be provided.")
        diff = _mp.pair_op(graph, "sub", coord_feat, coord_feat)                       # x_u - x_v
        radial = diff.square().sum(dim=1, keepdim=True)
        direction = diff / (radial.sqrt() + 1e-30)
        msg_h = self.edge_mlp[1:](self._first_edge_layer(graph, node_feat, radial, edge_feat))
        msg_x = self.coord_mlp(msg_h) * direction
        h = self.node_mlp(torch.cat([node_feat, _mp.edge_reduce(graph, msg_h, "sum")], dim=-1))
        return h, coord_feat + _mp.edge_reduce(graph, msg_x, "mean")


# =====================================================================================================================
#  TWIRLS
# =====================================================================================================================
class TWIRLSConv(nn.Module):
    r"""TWIRLS (graph networks from unfolded iteratively-reweighted least squares): MLP -> ``prop_step`` unfolded
    energy-descent steps with an optional IRLS attention re-weighting half-way -> MLP."""

    def __init__(self, input_d, output_d, hidden_d, prop_step, num_mlp_before=1, num_mlp_after=1, norm="none",
                 precond=True, alp=0, lam=1, attention=False, tau=0.2, T=-1, p=1, use_eta=False, attn_bef=False,
                 dropout=0.0, attn_dropout=0.0, inp_dropout=0.0):
        super().__init__()
        self.input_d, self.output_d, self.hidden_d, self.prop_step = input_d, output_d, hidden_d, prop_step
        self.num_mlp_before, self.num_mlp_after, self.norm, self.precond = num_mlp_before, num_mlp_after, norm, precond
        self.attention, self.alp, self.lam, self.tau, self.T, self.p = attention, alp, lam, tau, T, p
        self.use_eta, self.init_att = use_eta, attn_bef
        self.dropout, self.attn_dropout, self.inp_dropout = dropout, attn_dropout, inp_dropout
        self.attn_aft = prop_step // 2 if attention else -1
        # without learned pre-processing and attention the unfolding depends on the input only: compute it once
        self.cacheable = (not attention) and num_mlp_before == 0 and inp_dropout <= 0
        if self.cacheable:
            self.cached_unfolding = None
        width_in = input_d if num_mlp_before == 0 else hidden_d       # what mlp_aft receives
        width_mid = output_d if num_mlp_after == 0 else hidden_d      # what mlp_bef emits
        self.size_bef_unf, self.size_aft_unf = width_mid, width_in
        self.mlp_bef = MLP(input_d, hidden_d, width_mid, num_mlp_before, dropout, norm, init_activate=False)
        self.unfolding = TWIRLSUnfoldingAndAttention(hidden_d, alp, lam, prop_step, self.attn_aft, tau, T, p, use_eta,
                                                     attn_bef, attn_dropout, precond)
        self.mlp_aft = MLP(width_in, hidden_d, output_d, num_mlp_after, dropout, norm,
                           init_activate=num_mlp_before > 0 and num_mlp_after > 0)

    def forward(self, graph, feat):
        graph = graph.remove_self_loop().add_self_loop()
        if self.cacheable:
            if self.cached_unfolding is None:
                self.cached_unfolding = self.unfolding(graph, feat)
            mid = self.cached_unfolding
        else:
            x = F.dropout(feat, self.inp_dropout, training=self.training) if self.inp_dropout > 0 else feat
            mid = self.unfolding(graph, self.mlp_bef(x))
        return self.mlp_aft(mid)
