"""Dense-adjacency layers (DenseGraphConv / DenseSAGEConv / DenseChebConv), PNA / DGN and AtomicConv.

API parity: /root/reference/python/dgl/nn/pytorch/conv/{densegraphconv,densesageconv,densechebconv,pnaconv,dgnconv,
atomicconv}.py (values checked in tests/test_nn_reference_parity.py).

PNA / DGN do NOT use a degree-bucketed mailbox reduce (the reference's ``update_all`` with Python UDFs, one launch
sequence per distinct in-degree): every aggregator is a composition of segment reductions over the per-edge message
tensor - mean / sum / max / min directly, variance and higher moments through reductions of centred powers, the
directional aggregators of DGN through |F|-weighted sums - so a tower costs a fixed, small number of g-SpMM launches
whatever the degree distribution is. The message MLP ``M([h_u | h_v | e])`` is split per endpoint and applied to the
NODES (two N x d GEMMs + one per-edge add instead of an E x 2d GEMM on gathered features).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...utils import check_eq_shape
from .. import _mp


def _col(v, like):
    return v.view(-1, *([1] * (like.dim() - 1)))


# =====================================================================================================================
#  dense adjacency
# =====================================================================================================================
class DenseGraphConv(nn.Module):
    """GraphConv on a dense adjacency matrix ``adj[dst, src]``."""

    def __init__(self, in_feats, out_feats, norm="both", bias=True, activation=None):
        super().__init__()
        self._in_feats, self._out_feats, self._norm, self._activation = in_feats, out_feats, norm, activation
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, adj, feat):
        adj = adj.to(feat)
        x = feat
        if self._norm == "both":
            x = x * _col(adj.sum(0).clamp(min=1).rsqrt(), x)
        # GEMM on the narrower side of the propagation
        y = adj @ (x @ self.weight) if self._in_feats > self._out_feats else (adj @ x) @ self.weight
        if self._norm != "none":
            d_in = adj.sum(1).clamp(min=1)
            y = y * _col(d_in.rsqrt() if self._norm == "both" else d_in.reciprocal(), y)
        if self.bias is not None:
            y = y + self.bias
        return y if self._activation is None else self._activation(y)


class DenseSAGEConv(nn.Module):
    """GraphSAGE ('gcn' aggregator) on a dense adjacency matrix."""

    def __init__(self, in_feats, out_feats, feat_drop=0.0, bias=True, norm=None, activation=None):
        super().__init__()
        self._in_feats, self._out_feats, self._norm, self.activation = in_feats, out_feats, norm, activation
        self.feat_drop = nn.Dropout(feat_drop)
        self.fc = nn.Linear(in_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.fc.weight, gain=nn.init.calculate_gain("relu"))

    def forward(self, adj, feat):
        check_eq_shape(feat)
        pair = feat if isinstance(feat, tuple) else (feat, feat)
        x_src, x_dst = (self.feat_drop(t) for t in pair)
        adj = adj.to(x_src)
        pooled = torch.addmm(x_dst, adj, x_src) / (adj.sum(1, keepdim=True) + 1)     # (A x + x_self) / (deg + 1)
        out = self.fc(pooled)
        if self.activation is not None:
            out = self.activation(out)
        return out if self._norm is None else self._norm(out)


class DenseChebConv(nn.Module):
    """Chebyshev convolution on a dense adjacency matrix."""

    def __init__(self, in_feats, out_feats, k, bias=True):
        super().__init__()
        self._in_feats, self._out_feats, self._k = in_feats, out_feats, k
        self.W = nn.Parameter(torch.empty(k, in_feats, out_feats))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.bias is not None:
            nn.init.zeros_(self.bias)
        gain = nn.init.calculate_gain("relu")
        for w in self.W:
            nn.init.xavier_normal_(w, gain)

    def forward(self, adj, feat, lambda_max=None):
        A = adj.to(feat)
        s = A.sum(1).clamp(min=1).rsqrt()
        eye = torch.eye(A.shape[0], dtype=A.dtype, device=A.device)
        lap = eye - s.unsqueeze(1) * A * s.unsqueeze(0)
        if lambda_max is None:
            lambda_max = torch.linalg.eigvals(lap).real.max()
        lap_hat = 2 * lap / lambda_max - eye
        # T_i(L~) X by the three-term recurrence applied to the FEATURES (k propagations of [N, in]) instead of
        # materialising k dense [N, N] polynomials
        terms = [feat]
        if self._k > 1:
            terms.append(lap_hat @ feat)
        for _ in range(2, self._k):
            terms.append(2 * (lap_hat @ terms[-1]) - terms[-2])
        out = torch.einsum("kni,kio->no", torch.stack(terms, 0), self.W)
        return out if self.bias is None else out + self.bias


# =====================================================================================================================
#  PNA / DGN
# =====================================================================================================================
class _Neighbourhood:
    """Reductions of one tower's per-edge messages (computed lazily, shared between aggregators)."""

    def __init__(self, g, msg, h_dst=None, eig=None):
        self.g, self.msg, self.h_dst, self.eig = g, msg, h_dst, eig
        self._memo = {}

    def _get(self, key, make):
        if key not in self._memo:
            self._memo[key] = make()
        return self._memo[key]

    def reduce(self, how):
        return self._get(how, lambda: _mp.edge_reduce(self.g, self.msg, how))

    def centred_power_mean(self, n):
        centred = ops.gsddmm(self.g, "sub", self.msg, self.reduce("mean"), "e", "v")
        return _mp.edge_reduce(self.g, centred.pow(n), "mean")

    def variance(self):
        def make():
            mean = self.reduce("mean")
            return F.relu(_mp.edge_reduce(self.g, self.msg * self.msg, "mean") - mean * mean)
        return self._get("var", make)

    def field(self, k):
        """DGN vector field along eigenvector k: F_e = eig_k[src] - eig_k[dst], and its per-node L1 norm."""
        def make():
            col = self.eig[:, k:k + 1].contiguous()
            f = _mp.pair_op(self.g, "sub", col, col)
            return f, _mp.edge_reduce(self.g, f.abs(), "sum") + 1e-30
        return self._get(("field", k), make)


def _moment(n):
    def run(nb):
        m = nb.centred_power_mean(n)
        return torch.sign(m) * (m.abs() + 1e-30).pow(1.0 / n)
    return run


def _dir_av(k):
    def run(nb):
        f, norm = nb.field(k)
        return _mp.edge_reduce(nb.g, nb.msg * f.abs(), "sum") / norm
    return run


def _dir_dx(k):
    def run(nb):
        f, norm = nb.field(k)
        return (_mp.edge_reduce(nb.g, nb.msg * f, "sum") / norm - (_mp.edge_reduce(nb.g, f, "sum") / norm) * nb.h_dst).abs()
    return run


AGGREGATORS = {
    "mean": lambda nb: nb.reduce("mean"), "sum": lambda nb: nb.reduce("sum"),
    "max": lambda nb: nb.reduce("max"), "min": lambda nb: nb.reduce("min"),
    "var": lambda nb: nb.variance(), "std": lambda nb: (nb.variance() + 1e-30).sqrt(),
    "moment3": _moment(3), "moment4": _moment(4), "moment5": _moment(5),
}
for _k in range(1, 4):
    AGGREGATORS[f"dir{_k}-av"] = _dir_av(_k - 1)
    AGGREGATORS[f"dir{_k}-dx"] = _dir_dx(_k - 1)

SCALERS = {
    "identity": lambda h, logd, delta: h,
    "amplification": lambda h, logd, delta: h * (logd / delta),
    "attenuation": lambda h, logd, delta: h * (delta / logd),
}


class PNAConvTower(nn.Module):
    """One PNA tower: messages ``M([h_u | h_v | e])`` -> aggregators x degree scalers -> ``U([h | ...])`` -> BN."""

    def __init__(self, in_size, out_size, aggregators, scalers, delta, dropout=0.0, edge_feat_size=0):
        super().__init__()
        self.in_size, self.out_size, self.edge_feat_size = in_size, out_size, edge_feat_size
        self.aggregators, self.scalers, self.delta = aggregators, scalers, delta
        self.M = nn.Linear(2 * in_size + edge_feat_size, in_size)
        self.U = nn.Linear((len(aggregators) * len(scalers) + 1) * in_size, out_size)
        self.dropout = nn.Dropout(dropout)
        self.batchnorm = nn.BatchNorm1d(out_size)

    def _messages(self, graph, h, edge_feat):
        d, W = self.in_size, self.M.weight
        m = _mp.pair_op(graph, "add", F.linear(h, W[:, :d]), F.linear(h, W[:, d:2 * d])) + self.M.bias
        if self.edge_feat_size > 0:
            if edge_feat is None:
                raise AssertionError("Edge features must be provided.")
            m = m + F.linear(edge_feat, W[:, 2 * d:])
        return m

    def _pooled(self, graph, nb):
        has_nbr = (graph.in_degrees() > 0)
        logd = torch.log(graph.in_degrees().to(nb.msg.dtype).clamp(min=1) + 1).unsqueeze(-1)
        stats = torch.cat([AGGREGATORS[a](nb) for a in self.aggregators], dim=1)
        scaled = torch.cat([SCALERS[s](stats, logd, self.delta) for s in self.scalers], dim=1)
        return scaled * has_nbr.to(scaled.dtype).unsqueeze(-1)        # nodes without in-edges receive nothing

    def forward(self, graph, node_feat, edge_feat=None, eig_vec=None):
        nb = _Neighbourhood(graph, self._messages(graph, node_feat, edge_feat), node_feat, eig_vec)
        h = self.U(torch.cat([node_feat, self._pooled(graph, nb)], dim=-1))
        sizes = graph.batch_num_nodes().to(h.device)                   # graph-size normalisation 1 / sqrt(|V_graph|)
        h = h * torch.repeat_interleave(sizes.to(h.dtype).rsqrt(), sizes).unsqueeze(-1)
        return self.dropout(self.batchnorm(h))


class PNAConv(nn.Module):
    r"""Principal neighbourhood aggregation: ``num_towers`` towers on feature slices, mixed by a linear layer."""

    _tower = PNAConvTower

    def __init__(self, in_size, out_size, aggregators, scalers, delta, dropout=0.0, num_towers=1, edge_feat_size=0,
                 residual=True):
        super().__init__()
        if in_size % num_towers or out_size % num_towers:
            raise AssertionError("in_size and out_size must be divisible by num_towers")
        self.in_size, self.out_size, self.edge_feat_size = in_size, out_size, edge_feat_size
        self.tower_in_size, self.tower_out_size = in_size // num_towers, out_size // num_towers
        self.residual = residual and in_size == out_size
        self.towers = nn.ModuleList([self._tower(self.tower_in_size, self.tower_out_size, aggregators, scalers, delta,
                                                 dropout=dropout, edge_feat_size=edge_feat_size)
                                     for _ in range(num_towers)])
        self.mixing_layer = nn.Sequential(nn.Linear(out_size, out_size), nn.LeakyReLU())

    def forward(self, graph, node_feat, edge_feat=None, **tower_kw):
        slices = node_feat.split(self.tower_in_size, dim=1)
        mixed = self.mixing_layer(torch.cat([t(graph, x, edge_feat, **tower_kw) for t, x in zip(self.towers, slices)], 1))
        return mixed + node_feat if self.residual else mixed


class DGNConvTower(PNAConvTower):
    """PNA tower whose aggregator set also contains the directional ``dir{k}-av`` / ``dir{k}-dx`` aggregators."""


class DGNConv(PNAConv):
    r"""Directional graph network layer: PNA plus aggregators along the gradient field of Laplacian eigenvectors."""

    _tower = DGNConvTower

    def __init__(self, in_size, out_size, aggregators, scalers, delta, dropout=0.0, num_towers=1, edge_feat_size=0,
                 residual=True):
        super().__init__(in_size, out_size, aggregators, scalers, delta, dropout, num_towers, edge_feat_size, residual)
        self.use_eig_vec = any(a.startswith("dir") for a in aggregators)

    def forward(self, graph, node_feat, edge_feat=None, eig_vec=None):
        if self.use_eig_vec and eig_vec is None:
            raise AssertionError("Eigenvectors must be provided for the directional aggregators.")
        return super().forward(graph, node_feat, edge_feat, eig_vec=eig_vec)


# =====================================================================================================================
#  AtomicConv
# =====================================================================================================================
class RadialPooling(nn.Module):
    """K radial filters ``exp(-s_k (d - mu_k)^2) * cosine_cutoff_k(d)`` of the pair distances: [K, E, 1]."""

    def __init__(self, interaction_cutoffs, rbf_kernel_means, rbf_kernel_scaling):
        super().__init__()
        as_param = lambda t: nn.Parameter(t.reshape(-1, 1, 1), requires_grad=True)  # noqa: E731
        self.interaction_cutoffs = as_param(interaction_cutoffs)
        self.rbf_kernel_means = as_param(rbf_kernel_means)
        self.rbf_kernel_scaling = as_param(rbf_kernel_scaling)

    def forward(self, distances):
        rbf = torch.exp(-self.rbf_kernel_scaling * (distances - self.rbf_kernel_means).square())
        inside = distances <= self.interaction_cutoffs
        taper = 0.5 * (torch.cos(math.pi * distances / self.interaction_cutoffs) + 1)
        return rbf * torch.where(inside, taper, torch.zeros_like(taper))


class AtomicConv(nn.Module):
    r"""Atomic convolution (Gomes et al.): per radial filter k and atom type t, the filter-weighted count of
    neighbours of type t: ``out[i, k, t] = sum_j rbf_k(d_ij) [z_j == t]`` - one weighted g-SpMM."""

    def __init__(self, interaction_cutoffs, rbf_kernel_means, rbf_kernel_scaling, features_to_use=None):
        super().__init__()
        self.radial_pooling = RadialPooling(interaction_cutoffs, rbf_kernel_means, rbf_kernel_scaling)
        if features_to_use is None:
            self.num_channels, self.features_to_use = 1, None
        else:
            self.num_channels = len(features_to_use)
            self.features_to_use = nn.Parameter(features_to_use, requires_grad=False)

    def forward(self, graph, feat, distances):
        filt = self.radial_pooling(distances).squeeze(-1).t()                 # [E, K]
        if self.features_to_use is not None:
            feat = (feat == self.features_to_use).float()                      # one-hot over the selected atom types
        out = _mp.gather_sum(graph, feat.unsqueeze(1), filt.unsqueeze(-1))     # [N, K, T]
        return out.reshape(graph.num_nodes(), -1)


def msg_func(edges):
    """UDF form of the AtomicConv message (kept for API parity; the layer uses the fused weighted g-SpMM)."""
    return {"m": edges.src["hv"].unsqueeze(1) * edges.data["he"].unsqueeze(-1)}


def reduce_func(nodes):
    """UDF form of the AtomicConv reduction (kept for API parity)."""
    return {"hv_new": nodes.mailbox["m"].sum(1)}


# =====================================================================================================================
#  mailbox-form aggregators / scalers (public helpers of the reference's pnaconv / dgnconv modules). They act on a
#  dense mailbox ``h`` [nodes, degree, d] - the layers above never build one, these exist for user code that does.
# =====================================================================================================================
def aggregate_mean(h):
    return h.mean(dim=1)


def aggregate_sum(h):
    return h.sum(dim=1)


def aggregate_max(h):
    return h.amax(dim=1)


def aggregate_min(h):
    return h.amin(dim=1)


def aggregate_var(h):
    return F.relu((h * h).mean(dim=1) - h.mean(dim=1).square())


def aggregate_std(h):
    return (aggregate_var(h) + 1e-30).sqrt()


def _aggregate_moment(h, n):
    central = (h - h.mean(dim=1, keepdim=True)).pow(n).mean(dim=1)
    return torch.sign(central) * (central.abs() + 1e-30).pow(1.0 / n)


def aggregate_moment_3(h):
    return _aggregate_moment(h, 3)


def aggregate_moment_4(h):
    return _aggregate_moment(h, 4)


def aggregate_moment_5(h):
    return _aggregate_moment(h, 5)


def scale_identity(h, D=None, delta=None):
    return h


def scale_amplification(h, D, delta):
    return h * (math.log(D + 1) / delta)


def scale_attenuation(h, D, delta):
    return h * (delta / math.log(D + 1))


def _mailbox_field(eig_s, eig_d, eig_idx):
    f = eig_s[:, :, eig_idx] - eig_d[:, :, eig_idx]
    return (f / (f.abs().sum(dim=1, keepdim=True) + 1e-30)).unsqueeze(-1)


def aggregate_dir_av(h, eig_s, eig_d, eig_idx):
    return (h * _mailbox_field(eig_s, eig_d, eig_idx).abs()).sum(dim=1)


def aggregate_dir_dx(h, eig_s, eig_d, h_in, eig_idx):
    w = _mailbox_field(eig_s, eig_d, eig_idx)
    return ((h * w).sum(dim=1) - w.sum(dim=1) * h_in).abs()
