"""Dense, PNA/DGN, AtomicConv layers (parity: /root/reference/python/dgl/nn/pytorch/conv/
{densegraphconv,densesageconv,densechebconv,pnaconv,dgnconv,atomicconv}.py)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import function as fn
from ...base import DGLError
from ...utils import check_eq_shape


class DenseGraphConv(nn.Module):
    """GraphConv on a dense adjacency matrix."""

    def __init__(self, in_feats, out_feats, norm="both", bias=True, activation=None):
        super().__init__()
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()
        self._activation = activation

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, adj, feat):
        adj = adj.to(feat)
        src_degrees = adj.sum(dim=0).clamp(min=1)
        dst_degrees = adj.sum(dim=1).clamp(min=1)
        feat_src = feat
        if self._norm == "both":
            norm_src = torch.pow(src_degrees, -0.5)
            feat_src = feat_src * norm_src.view(-1, *([1] * (feat.dim() - 1)))
        if self._in_feats > self._out_feats:
            rst = adj @ torch.matmul(feat_src, self.weight)
        else:
            rst = torch.matmul(adj @ feat_src, self.weight)
        if self._norm != "none":
            norm_dst = torch.pow(dst_degrees, -0.5) if self._norm == "both" else 1.0 / dst_degrees
            rst = rst * norm_dst.view(-1, *([1] * (rst.dim() - 1)))
        if self.bias is not None:
            rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst


class DenseSAGEConv(nn.Module):
    """GraphSAGE-gcn on a dense adjacency matrix."""

    def __init__(self, in_feats, out_feats, feat_drop=0.0, bias=True, norm=None, activation=None):
        super().__init__()
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self.feat_drop = nn.Dropout(feat_drop)
        self.activation = activation
        self.fc = nn.Linear(in_feats, out_feats, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.fc.weight, gain=nn.init.calculate_gain("relu"))

    def forward(self, adj, feat):
        check_eq_shape(feat)
        if isinstance(feat, tuple):
            feat_src = self.feat_drop(feat[0])
            feat_dst = self.feat_drop(feat[1])
        else:
            feat_src = feat_dst = self.feat_drop(feat)
        adj = adj.to(feat_src)
        in_degrees = adj.sum(dim=1, keepdim=True)
        h_neigh = (adj @ feat_src + feat_dst) / (in_degrees + 1)
        rst = self.fc(h_neigh)
        if self.activation is not None:
            rst = self.activation(rst)
        if self._norm is not None:
            rst = self._norm(rst)
        return rst


class DenseChebConv(nn.Module):
    """ChebConv on a dense adjacency matrix."""

    def __init__(self, in_feats, out_feats, k, bias=True):
        super().__init__()
        self._in_feats, self._out_feats, self._k = in_feats, out_feats, k
        self.W = nn.Parameter(torch.Tensor(k, in_feats, out_feats))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_feats))
        else:
            self.register_buffer("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.bias is not None:
            nn.init.zeros_(self.bias)
        for i in range(self._k):
            nn.init.xavier_normal_(self.W[i], nn.init.calculate_gain("relu"))

    def forward(self, adj, feat, lambda_max=None):
        A = adj.to(feat)
        num_nodes = A.shape[0]
        in_degree = 1 / A.sum(dim=1).clamp(min=1).sqrt()
        D_invsqrt = torch.diag(in_degree)
        I = torch.eye(num_nodes).to(A)
        L = I - D_invsqrt @ A @ D_invsqrt
        if lambda_max is None:
            lambda_ = torch.linalg.eigvals(L).real
            lambda_max = lambda_.max()
        L_hat = 2 * L / lambda_max - I
        Z = [torch.eye(num_nodes).to(A)]
        for i in range(1, self._k):
            if i == 1:
                Z.append(L_hat)
            else:
                Z.append(2 * L_hat @ Z[-1] - Z[-2])
        Zs = torch.stack(Z, 0)
        Zh = Zs @ feat.unsqueeze(0) @ self.W
        Zh = Zh.sum(0)
        if self.bias is not None:
            Zh = Zh + self.bias
        return Zh


# ----------------------------------------------------------------------------- PNA / DGN
def aggregate_mean(h):
    return torch.mean(h, dim=1)


def aggregate_max(h):
    return torch.max(h, dim=1)[0]


def aggregate_min(h):
    return torch.min(h, dim=1)[0]


def aggregate_sum(h):
    return torch.sum(h, dim=1)


def aggregate_std(h):
    return torch.sqrt(aggregate_var(h) + 1e-30)


def aggregate_var(h):
    h_mean_squares = torch.mean(h * h, dim=1)
    h_mean = torch.mean(h, dim=1)
    return torch.relu(h_mean_squares - h_mean * h_mean)


def _aggregate_moment(h, n):
    h_mean = torch.mean(h, dim=1, keepdim=True)
    h_n = torch.mean(torch.pow(h - h_mean, n), dim=1)
    rooted = torch.sign(h_n) * torch.pow(torch.abs(h_n) + 1e-30, 1.0 / n)
    return rooted


def aggregate_moment_3(h):
    return _aggregate_moment(h, n=3)


def aggregate_moment_4(h):
    return _aggregate_moment(h, n=4)


def aggregate_moment_5(h):
    return _aggregate_moment(h, n=5)


def scale_identity(h, D=None, delta=None):
    return h


def scale_amplification(h, D, delta):
    return h * (np.log(D + 1) / delta)


def scale_attenuation(h, D, delta):
    return h * (delta / np.log(D + 1))


AGGREGATORS = {"mean": aggregate_mean, "sum": aggregate_sum, "max": aggregate_max, "min": aggregate_min,
               "std": aggregate_std, "var": aggregate_var, "moment3": aggregate_moment_3,
               "moment4": aggregate_moment_4, "moment5": aggregate_moment_5}
SCALERS = {"identity": scale_identity, "amplification": scale_amplification, "attenuation": scale_attenuation}


class PNAConvTower(nn.Module):
    def __init__(self, in_size, out_size, aggregators, scalers, delta, dropout=0.0, edge_feat_size=0):
        super().__init__()
        self.in_size, self.out_size = in_size, out_size
        self.aggregators, self.scalers, self.delta = aggregators, scalers, delta
        self.edge_feat_size = edge_feat_size
        self.M = nn.Linear(2 * in_size + edge_feat_size, in_size)
        self.U = nn.Linear((len(aggregators) * len(scalers) + 1) * in_size, out_size)
        self.dropout = nn.Dropout(dropout)
        self.batchnorm = nn.BatchNorm1d(out_size)

    def reduce_func(self, nodes):
        msg = nodes.mailbox["msg"]
        degree = msg.size(1)
        h = torch.cat([AGGREGATORS[agg](msg) for agg in self.aggregators], dim=1)
        h = torch.cat([SCALERS[scaler](h, D=degree, delta=self.delta) if scaler != "identity" else h
                       for scaler in self.scalers], dim=1)
        return {"h_neigh": h}

    def message(self, edges):
        if self.edge_feat_size > 0:
            f = torch.cat([edges.src["h"], edges.dst["h"], edges.data["a"]], dim=-1)
        else:
            f = torch.cat([edges.src["h"], edges.dst["h"]], dim=-1)
        return {"msg": self.M(f)}

    def forward(self, graph, node_feat, edge_feat=None):
        with graph.local_scope():
            graph.ndata["h"] = node_feat
            if self.edge_feat_size > 0:
                assert edge_feat is not None, "Edge features must be provided."
                graph.edata["a"] = edge_feat
            graph.update_all(self.message, self.reduce_func)
            h = self.U(torch.cat([node_feat, graph.ndata["h_neigh"]], dim=-1))
            h = h * (graph.in_degrees() > 0).to(h).unsqueeze(-1) if False else h
            h = self.dropout(h)
            return self.batchnorm(h)


class PNAConv(nn.Module):
    r"""Principal neighbourhood aggregation (reference pnaconv.py): multiple aggregators x degree scalers."""

    def __init__(self, in_size, out_size, aggregators, scalers, delta, dropout=0.0, num_towers=1, edge_feat_size=0,
                 residual=True):
        super().__init__()
        self.in_size, self.out_size = in_size, out_size
        assert in_size % num_towers == 0, "in_size must be divisible by num_towers"
        assert out_size % num_towers == 0, "out_size must be divisible by num_towers"
        self.tower_in_size = in_size // num_towers
        self.tower_out_size = out_size // num_towers
        self.edge_feat_size = edge_feat_size
        self.residual = residual
        if self.in_size != self.out_size:
            self.residual = False
        self.towers = nn.ModuleList([PNAConvTower(self.tower_in_size, self.tower_out_size, aggregators, scalers, delta,
                                                  dropout=dropout, edge_feat_size=edge_feat_size)
                                     for _ in range(num_towers)])
        self.mixing_layer = nn.Sequential(nn.Linear(out_size, out_size), nn.LeakyReLU())

    def forward(self, graph, node_feat, edge_feat=None):
        h_cat = torch.cat([tower(graph, node_feat[:, ti * self.tower_in_size: (ti + 1) * self.tower_in_size], edge_feat)
                           for ti, tower in enumerate(self.towers)], dim=1)
        h_out = self.mixing_layer(h_cat)
        if self.residual:
            h_out = h_out + node_feat
        return h_out


def aggregate_dir_av(h, eig_s, eig_d, eig_idx):
    h_mod = torch.mul(h, (torch.abs(eig_s[:, :, eig_idx] - eig_d[:, :, eig_idx]) /
                          (torch.sum(torch.abs(eig_s[:, :, eig_idx] - eig_d[:, :, eig_idx]), keepdim=True, dim=1) + 1e-30)).unsqueeze(-1))
    return torch.sum(h_mod, dim=1)


def aggregate_dir_dx(h, eig_s, eig_d, h_in, eig_idx):
    eig_w = ((eig_s[:, :, eig_idx] - eig_d[:, :, eig_idx]) /
             (torch.sum(torch.abs(eig_s[:, :, eig_idx] - eig_d[:, :, eig_idx]), keepdim=True, dim=1) + 1e-30)).unsqueeze(-1)
    h_mod = torch.mul(h, eig_w)
    return torch.abs(torch.sum(h_mod, dim=1) - torch.sum(eig_w, dim=1) * h_in)


class DGNConvTower(PNAConvTower):
    def message(self, edges):
        if self.edge_feat_size > 0:
            f = torch.cat([edges.src["h"], edges.dst["h"], edges.data["a"]], dim=-1)
        else:
            f = torch.cat([edges.src["h"], edges.dst["h"]], dim=-1)
        return {"msg": self.M(f), "eig_s": edges.src["eig"], "eig_d": edges.dst["eig"]}

    def reduce_func(self, nodes):
        h_in = nodes.data["h"]
        eig_s = nodes.mailbox["eig_s"]
        eig_d = nodes.mailbox["eig_d"]
        msg = nodes.mailbox["msg"]
        degree = msg.size(1)
        outs = []
        for agg in self.aggregators:
            if agg.startswith("dir"):
                k = int(agg[3: agg.index("-")]) if "-" in agg else 1
                kind = agg.split("-")[1]
                if kind == "av":
                    outs.append(aggregate_dir_av(msg, eig_s, eig_d, k - 1))
                else:
                    outs.append(aggregate_dir_dx(msg, eig_s, eig_d, h_in, k - 1))
            else:
                outs.append(AGGREGATORS[agg](msg))
        h = torch.cat(outs, dim=1)
        h = torch.cat([SCALERS[scaler](h, D=degree, delta=self.delta) if scaler != "identity" else h
                       for scaler in self.scalers], dim=1)
        return {"h_neigh": h}

    def forward(self, graph, node_feat, edge_feat=None, eig_vec=None):
        with graph.local_scope():
            graph.ndata["h"] = node_feat
            if eig_vec is not None:
                graph.ndata["eig"] = eig_vec
            else:
                graph.ndata["eig"] = node_feat.new_zeros((node_feat.shape[0], 1))
            if self.edge_feat_size > 0:
                graph.edata["a"] = edge_feat
            graph.update_all(self.message, self.reduce_func)
            h = self.U(torch.cat([node_feat, graph.ndata["h_neigh"]], dim=-1))
            h = self.dropout(h)
            return self.batchnorm(h)


class DGNConv(PNAConv):
    r"""Directional graph network layer (reference dgnconv.py): PNA plus directional aggregators
    ``dir{k}-av`` / ``dir{k}-dx`` driven by Laplacian eigenvectors."""

    def __init__(self, in_size, out_size, aggregators, scalers, delta, dropout=0.0, num_towers=1, edge_feat_size=0,
                 residual=True):
        super().__init__(in_size, out_size, aggregators, scalers, delta, dropout, num_towers, edge_feat_size, residual)
        self.towers = nn.ModuleList([DGNConvTower(self.tower_in_size, self.tower_out_size, aggregators, scalers, delta,
                                                  dropout=dropout, edge_feat_size=edge_feat_size)
                                     for _ in range(num_towers)])
        self.use_eig_vec = any(a.startswith("dir") for a in aggregators)

    def forward(self, graph, node_feat, edge_feat=None, eig_vec=None):
        h_cat = torch.cat([tower(graph, node_feat[:, ti * self.tower_in_size: (ti + 1) * self.tower_in_size], edge_feat,
                                 eig_vec) for ti, tower in enumerate(self.towers)], dim=1)
        h_out = self.mixing_layer(h_cat)
        if self.residual:
            h_out = h_out + node_feat
        return h_out


# ----------------------------------------------------------------------------- AtomicConv
def msg_func(edges):
    return {"ev": edges.data["he"], "hv": edges.src["hv"], "uv": edges.src["hv"] * 0 + edges.dst["hv"]}


def reduce_func(nodes):
    """UDF counterpart of the AtomicConv aggregation: sum the per-edge radial messages of each node (the layer itself
    uses the fused ``u_mul_e`` + ``sum`` kernel)."""
    return {"hv_new": nodes.mailbox["m"].sum(1)}


class RadialPooling(nn.Module):
    def __init__(self, interaction_cutoffs, rbf_kernel_means, rbf_kernel_scaling):
        super().__init__()
        self.interaction_cutoffs = nn.Parameter(interaction_cutoffs.reshape(-1, 1, 1), requires_grad=True)
        self.rbf_kernel_means = nn.Parameter(rbf_kernel_means.reshape(-1, 1, 1), requires_grad=True)
        self.rbf_kernel_scaling = nn.Parameter(rbf_kernel_scaling.reshape(-1, 1, 1), requires_grad=True)

    def forward(self, distances):
        scaled_euclidean_distance = -self.rbf_kernel_scaling * (distances - self.rbf_kernel_means) ** 2
        rbf_kernel_results = torch.exp(scaled_euclidean_distance)
        cos_values = 0.5 * (torch.cos(np.pi * distances / self.interaction_cutoffs) + 1)
        cutoff_values = torch.where(distances <= self.interaction_cutoffs, cos_values, torch.zeros_like(cos_values))
        return rbf_kernel_results * cutoff_values


class AtomicConv(nn.Module):
    r"""Atomic convolution layer (reference atomicconv.py)."""

    def __init__(self, interaction_cutoffs, rbf_kernel_means, rbf_kernel_scaling, features_to_use=None):
        super().__init__()
        self.radial_pooling = RadialPooling(interaction_cutoffs, rbf_kernel_means, rbf_kernel_scaling)
        if features_to_use is None:
            self.num_channels = 1
            self.features_to_use = None
        else:
            self.num_channels = len(features_to_use)
            self.features_to_use = nn.Parameter(features_to_use, requires_grad=False)

    def forward(self, graph, feat, distances):
        with graph.local_scope():
            radial_pooled_values = self.radial_pooling(distances)  # (K, E, 1)
            if self.features_to_use is not None:
                feat = (feat == self.features_to_use).float()  # (V, T)
            graph.ndata["hv"] = feat
            graph.edata["he"] = radial_pooled_values.transpose(1, 0).squeeze(-1)  # (E, K)
            # m = he (E,K,1) * hv_src (E,1,T)
            graph.edata["he3"] = graph.edata["he"].unsqueeze(-1)
            graph.srcdata["hv3"] = feat.unsqueeze(1)
            graph.update_all(fn.u_mul_e("hv3", "he3", "m"), fn.sum("m", "hv_new"))
            return graph.ndata["hv_new"].view(graph.num_nodes(), -1)
