"""Graph-transformer building blocks: BiasedMHA, GraphormerLayer, EGTLayer and the structural encoders (degree, path,
spatial, 3-d spatial, Laplacian positional).

API parity: /root/reference/python/dgl/nn/pytorch/gt/*.py (constructor arguments and parameter names; values are
checked against the unmodified reference in tests/test_nn_reference_parity.py). These modules work on dense padded
batches [batch, N, d] by definition; attention is written head-major with einsum contractions (one batched GEMM for
QK^T and one for the value mix, no reshapes through [N, batch * heads] layouts); fully masked rows yield zeros, not NaN.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["BiasedMHA", "BiasedMultiheadAttention", "GraphormerLayer", "EGTLayer", "DegreeEncoder", "PathEncoder",
           "SpatialEncoder", "SpatialEncoder3d", "LapPosEncoder"]


def gaussian(x, mean, std):
    """Gaussian basis function with the truncated constant pi = 3.14159 used by the reference encoders."""
    return torch.exp(-0.5 * ((x - mean) / std) ** 2) / ((2 * 3.14159) ** 0.5 * std)


class BiasedMHA(nn.Module):
    """Dense multi-head self-attention with a per-pair, per-head bias ('add' to or 'mul' with the scores) and a padding /
    attention mask (True = excluded)."""

    def __init__(self, feat_size, num_heads, bias=True, attn_bias_type="add", attn_drop=0.1):
        super().__init__()
        if feat_size % num_heads:
            raise AssertionError("feat_size must be divisible by num_heads")
        self.feat_size, self.num_heads, self.head_dim = feat_size, num_heads, feat_size // num_heads
        self.scaling, self.attn_bias_type = self.head_dim ** -0.5, attn_bias_type
        for name in ("q_proj", "k_proj", "v_proj", "out_proj"):
            setattr(self, name, nn.Linear(feat_size, feat_size, bias=bias))
        self.dropout = nn.Dropout(p=attn_drop)
        self.reset_parameters()

    def reset_parameters(self):
        for lin in (self.q_proj, self.k_proj, self.v_proj):
            nn.init.xavier_uniform_(lin.weight, gain=2 ** -0.5)
        nn.init.xavier_uniform_(self.out_proj.weight)
        if self.out_proj.bias is not None:
            nn.init.zeros_(self.out_proj.bias)

    def _heads(self, x):
        b, n, _ = x.shape
        return x.view(b, n, self.num_heads, self.head_dim).transpose(1, 2)        # [B, H, N, d]

    def forward(self, ndata, attn_bias=None, attn_mask=None):
        q, k, v = self._heads(self.q_proj(ndata)), self._heads(self.k_proj(ndata)), self._heads(self.v_proj(ndata))
        bias = None if attn_bias is None else attn_bias.permute(0, 3, 1, 2)        # [B, N, N, H] -> [B, H, N, N]
        mask = None
        if attn_mask is not None:
            mask = attn_mask.to(torch.bool)
            mask = mask.unsqueeze(1) if mask.dim() == 3 else mask.permute(0, 3, 1, 2)
        score = torch.einsum("bhid,bhjd->bhij", q * self.scaling, k)
        if bias is not None:
            score = score + bias if self.attn_bias_type == "add" else score * bias
        if mask is not None:
            score = score.masked_fill(mask, float("-inf"))
        prob = self.dropout(torch.nan_to_num(score.softmax(dim=-1)))               # fully masked rows -> zeros
        mixed = torch.einsum("bhij,bhjd->bihd", prob, v)
        return self.out_proj(mixed.reshape(ndata.shape[0], ndata.shape[1], self.feat_size))


BiasedMultiheadAttention = BiasedMHA


class GraphormerLayer(nn.Module):
    """Transformer encoder layer (pre- or post-norm) whose attention is :class:`BiasedMHA`."""

    def __init__(self, feat_size, hidden_size, num_heads, attn_bias_type="add", norm_first=False, dropout=0.1,
                 attn_dropout=0.1, activation=nn.ReLU()):
        super().__init__()
        self.norm_first = norm_first
        self.attn = BiasedMHA(feat_size=feat_size, num_heads=num_heads, attn_bias_type=attn_bias_type,
                              attn_drop=attn_dropout)
        self.ffn = nn.Sequential(nn.Linear(feat_size, hidden_size), activation, nn.Dropout(p=dropout),
                                 nn.Linear(hidden_size, feat_size), nn.Dropout(p=dropout))
        self.dropout = nn.Dropout(p=dropout)
        self.attn_layer_norm = nn.LayerNorm(feat_size)
        self.ffn_layer_norm = nn.LayerNorm(feat_size)

    def _sublayer(self, x, norm, fn):
        if self.norm_first:
            return x + fn(norm(x))
        return norm(x + fn(x))

    def forward(self, nfeat, attn_bias=None, attn_mask=None):
        x = self._sublayer(nfeat, self.attn_layer_norm, lambda t: self.dropout(self.attn(t, attn_bias, attn_mask)))
        return self._sublayer(x, self.ffn_layer_norm, self.ffn)


class EGTLayer(nn.Module):
    """Edge-augmented graph transformer layer: pair (edge) channels bias and gate the attention and are updated from the
    clipped pre-softmax scores."""

    def __init__(self, feat_size, edge_feat_size, num_heads, num_virtual_nodes, dropout=0, attn_dropout=0,
                 activation=nn.ELU(), edge_update=True):
        super().__init__()
        if feat_size % num_heads:
            raise AssertionError("feat_size must be divisible by num_heads")
        self.num_heads, self.num_virtual_nodes, self.edge_update = num_heads, num_virtual_nodes, edge_update
        self.dot_dim = feat_size // num_heads
        self.mha_ln_h = nn.LayerNorm(feat_size)
        self.mha_ln_e = nn.LayerNorm(edge_feat_size)
        self.edge_input = nn.Linear(edge_feat_size, num_heads)
        self.qkv_proj = nn.Linear(feat_size, feat_size * 3)
        self.gate = nn.Linear(edge_feat_size, num_heads)
        self.attn_dropout = nn.Dropout(attn_dropout)
        self.node_output = nn.Linear(feat_size, feat_size)
        self.mha_dropout_h = nn.Dropout(dropout)

        def ffn(width):
            return nn.Sequential(nn.LayerNorm(width), nn.Linear(width, width), activation, nn.Linear(width, width),
                                 nn.Dropout(dropout))

        self.node_ffn = ffn(feat_size)
        if edge_update:
            self.edge_output = nn.Linear(num_heads, edge_feat_size)
            self.mha_dropout_e = nn.Dropout(dropout)
            self.edge_ffn = ffn(edge_feat_size)

    def forward(self, nfeat, efeat, mask=None):
        B, N, _ = nfeat.shape
        e_norm = self.mha_ln_e(efeat)
        q, k, v = self.qkv_proj(self.mha_ln_h(nfeat)).view(B, N, -1, self.num_heads).split(self.dot_dim, dim=2)
        raw = torch.einsum("bldh,bmdh->blmh", q, k).clamp(-5, 5) + self.edge_input(e_norm)   # [B, N, N, H]
        gate_logit = self.gate(e_norm)
        if mask is not None:
            pad = mask.unsqueeze(-1)
            gate, attn = torch.sigmoid(gate_logit + pad), (raw + pad).softmax(dim=2)
        else:
            gate, attn = torch.sigmoid(gate_logit), raw.softmax(dim=2)
        mixed = torch.einsum("blmh,bmkh->blkh", self.attn_dropout(attn * gate), v)
        scale = torch.log1p(gate.sum(dim=2, keepdim=True))          # degree scaler from the soft in-degree
        scale = torch.cat([torch.ones_like(scale[:, : self.num_virtual_nodes]), scale[:, self.num_virtual_nodes:]], 1)
        h = self.mha_dropout_h(self.node_output((mixed * scale).reshape(B, N, -1))) + nfeat
        h = h + self.node_ffn(h)
        if not self.edge_update:
            return h
        e = self.mha_dropout_e(self.edge_output(raw)) + efeat
        return h, e + self.edge_ffn(e)


class DegreeEncoder(nn.Module):
    """Learned embedding of (clamped) node degrees; 'both' sums an in-degree and an out-degree table."""

    def __init__(self, max_degree, embedding_dim, direction="both"):
        super().__init__()
        self.direction, self.max_degree = direction, max_degree
        if direction == "both":
            self.encoder1 = nn.Embedding(max_degree + 1, embedding_dim, padding_idx=0)
            self.encoder2 = nn.Embedding(max_degree + 1, embedding_dim, padding_idx=0)
        else:
            self.encoder = nn.Embedding(max_degree + 1, embedding_dim, padding_idx=0)

    def forward(self, degrees):
        d = degrees.clamp(min=0, max=self.max_degree)
        if self.direction == "both":
            return self.encoder1(d[0]) + self.encoder2(d[1])
        if self.direction in ("in", "out"):
            return self.encoder(d)
        raise ValueError(f'Supported direction options: "in", "out" and "both", but got {self.direction}')


class PathEncoder(nn.Module):
    """Attention bias from the edge features along shortest paths: the mean over path positions of <edge feature,
    learned per-position per-head vector>."""

    def __init__(self, max_len, feat_dim, num_heads=1):
        super().__init__()
        self.max_len, self.feat_dim, self.num_heads = max_len, feat_dim, num_heads
        self.embedding_table = nn.Embedding(max_len * num_heads, feat_dim)

    def forward(self, dist, path_data):
        steps = dist.clamp(min=1, max=self.max_len).unsqueeze(-1)
        table = self.embedding_table.weight.view(self.max_len, self.num_heads, self.feat_dim)
        return torch.einsum("bxyld,lhd->bxyh", path_data, table) / steps


class SpatialEncoder(nn.Module):
    """Attention bias looked up from the shortest-path distance (-1 = unreachable, padding index 0 after the shift)."""

    def __init__(self, max_dist, num_heads=1):
        super().__init__()
        self.max_dist, self.num_heads = max_dist, num_heads
        self.embedding_table = nn.Embedding(max_dist + 2, num_heads, padding_idx=0)

    def forward(self, dist):
        return self.embedding_table(dist.clamp(min=-1, max=self.max_dist) + 1)


class SpatialEncoder3d(nn.Module):
    """Attention bias from 3-d Euclidean distances: per node-type-pair affine map (gamma * d + beta, one learned scalar
    per source type plus one per target type) -> Gaussian basis -> 2-layer MLP."""

    def __init__(self, num_kernels, num_heads=1, max_node_type=100):
        super().__init__()
        self.num_kernels, self.num_heads, self.max_node_type = num_kernels, num_heads, max_node_type
        self.means = nn.Parameter(torch.empty(num_kernels).uniform_(0, 3))
        self.stds = nn.Parameter(torch.empty(num_kernels).uniform_(0, 3))
        self.linear_layer_1 = nn.Linear(num_kernels, num_kernels)
        self.linear_layer_2 = nn.Linear(num_kernels, num_heads)
        # row 0: no type given; rows 1 .. T+1: source types (1 = padding); rows T+2 .. 2T+2: target types
        self.gamma = nn.Embedding(2 * max_node_type + 3, 1, padding_idx=0)
        self.beta = nn.Embedding(2 * max_node_type + 3, 1, padding_idx=0)
        nn.init.constant_(self.gamma.weight, 1)
        nn.init.constant_(self.beta.weight, 0)

    def forward(self, coord, node_type=None):
        B, n = coord.shape[:2]
        dist = torch.cdist(coord, coord, p=2.0).unsqueeze(-1)           # [B, n, n, 1]
        if node_type is None:
            pair = torch.zeros(B, n, n, 2, dtype=torch.long, device=coord.device)
        else:
            pair = torch.stack([(node_type + 2).unsqueeze(-1).expand(-1, -1, n),
                                (node_type + self.max_node_type + 3).unsqueeze(1).expand(-1, n, -1)], dim=-1)
        scaled = self.gamma(pair).sum(dim=-2) * dist + self.beta(pair).sum(dim=-2)
        basis = gaussian(scaled.expand(-1, -1, -1, self.num_kernels), self.means, self.stds.abs() + 1e-2)
        return self.linear_layer_2(F.gelu(self.linear_layer_1(basis)))


def _relu_mlp(widths):
    layers = []
    for a, b in zip(widths[:-1], widths[1:]):
        layers += [nn.Linear(a, b), nn.ReLU()]
    return layers


class LapPosEncoder(nn.Module):
    """Learned Laplacian positional encoding: (eigenvector entry, eigenvalue) pairs of a node are encoded as a set by a
    Transformer or a DeepSet and summed."""

    def __init__(self, model_type, num_layer, k, dim, n_head=1, batch_norm=False, num_post_layer=0):
        super().__init__()
        self.model_type, self.k = model_type, k
        if model_type == "Transformer":
            self.linear = nn.Linear(2, dim)
            self.pe_encoder = nn.TransformerEncoder(nn.TransformerEncoderLayer(d_model=dim, nhead=n_head, batch_first=True),
                                                    num_layers=num_layer)
        elif model_type == "DeepSet":
            if num_layer == 1:
                body = [nn.ReLU()]
            else:
                self.linear = nn.Linear(2, 2 * dim)
                body = [nn.ReLU()] + _relu_mlp([2 * dim] * (num_layer - 1) + [dim])
            self.pe_encoder = nn.Sequential(*body)
        else:
            raise ValueError(f"model_type should be one of ['Transformer', 'DeepSet'], but got {model_type}")
        self.raw_norm = nn.BatchNorm1d(k) if batch_norm else None
        if num_post_layer > 0:
            mids = [dim, dim] if num_post_layer == 1 else [dim] + [2 * dim] * (num_post_layer - 1) + [dim]
            self.post_mlp = nn.Sequential(*_relu_mlp(mids))
        else:
            self.post_mlp = None

    def forward(self, eigvals, eigvecs):
        pe = torch.stack((eigvecs, eigvals), dim=2).float()             # [N, k, 2]
        missing = torch.isnan(pe)
        pe = pe.masked_fill(missing, 0.0)
        if self.raw_norm is not None:
            pe = self.raw_norm(pe)
        pe = self.linear(pe)
        if self.model_type == "Transformer":
            pe = self.pe_encoder(src=pe, src_key_padding_mask=missing[:, :, 1])
        else:
            pe = self.pe_encoder(pe)
        pe = pe.masked_fill(missing[:, :, 1].unsqueeze(-1), 0.0).sum(dim=1)
        return pe if self.post_mlp is None else self.post_mlp(pe)
