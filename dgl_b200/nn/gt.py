"""Graph-transformer building blocks (parity: /root/reference/python/dgl/nn/pytorch/gt/*.py): BiasedMHA,
GraphormerLayer, EGTLayer, DegreeEncoder, PathEncoder, SpatialEncoder, SpatialEncoder3d, LapPosEncoder.
Dense per-graph attention on padded batches, as in the reference."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["BiasedMHA", "BiasedMultiheadAttention", "GraphormerLayer", "EGTLayer", "DegreeEncoder", "PathEncoder",
           "SpatialEncoder", "SpatialEncoder3d", "LapPosEncoder"]


def gaussian(x, mean, std):
    """Gaussian basis function ``N(x; mean, std)`` (reference gt/spatial_encoder.py:8, constant pi = 3.14159)."""
    return torch.exp(-0.5 * ((x - mean) / std) ** 2) / ((2 * 3.14159) ** 0.5 * std)


class BiasedMHA(nn.Module):
    def __init__(self, feat_size, num_heads, bias=True, attn_bias_type="add", attn_drop=0.1):
        super().__init__()
        self.feat_size, self.num_heads = feat_size, num_heads
        self.head_dim = feat_size // num_heads
        assert self.head_dim * num_heads == feat_size, "feat_size must be divisible by num_heads"
        self.scaling = self.head_dim ** -0.5
        self.attn_bias_type = attn_bias_type
        self.q_proj = nn.Linear(feat_size, feat_size, bias=bias)
        self.k_proj = nn.Linear(feat_size, feat_size, bias=bias)
        self.v_proj = nn.Linear(feat_size, feat_size, bias=bias)
        self.out_proj = nn.Linear(feat_size, feat_size, bias=bias)
        self.dropout = nn.Dropout(p=attn_drop)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.q_proj.weight, gain=2 ** -0.5)
        nn.init.xavier_uniform_(self.k_proj.weight, gain=2 ** -0.5)
        nn.init.xavier_uniform_(self.v_proj.weight, gain=2 ** -0.5)
        nn.init.xavier_uniform_(self.out_proj.weight)
        if self.out_proj.bias is not None:
            nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, ndata, attn_bias=None, attn_mask=None):
        q_h = self.q_proj(ndata).transpose(0, 1)
        k_h = self.k_proj(ndata).transpose(0, 1)
        v_h = self.v_proj(ndata).transpose(0, 1)
        bsz, N, _ = ndata.shape
        q_h = q_h.reshape(N, bsz * self.num_heads, self.head_dim).transpose(0, 1) * self.scaling
        k_h = k_h.reshape(N, bsz * self.num_heads, self.head_dim).permute(1, 2, 0)
        v_h = v_h.reshape(N, bsz * self.num_heads, self.head_dim).transpose(0, 1)
        attn_weights = torch.bmm(q_h, k_h).transpose(0, 2).reshape(N, N, bsz, self.num_heads).transpose(0, 2)
        if attn_bias is not None:
            if self.attn_bias_type == "add":
                attn_weights = attn_weights + attn_bias
            else:
                attn_weights = attn_weights * attn_bias
        if attn_mask is not None:
            attn_weights = attn_weights.masked_fill(attn_mask.to(torch.bool).unsqueeze(-1) if attn_mask.dim() == 3 else attn_mask.to(torch.bool), float("-inf"))
        attn_weights = F.softmax(attn_weights.transpose(0, 2).reshape(N, N, bsz * self.num_heads).transpose(0, 2), dim=2)
        attn_weights = torch.nan_to_num(attn_weights)
        attn_weights = self.dropout(attn_weights)
        attn = torch.bmm(attn_weights, v_h).transpose(0, 1)
        attn = self.out_proj(attn.reshape(N, bsz, self.feat_size).transpose(0, 1))
        return attn


BiasedMultiheadAttention = BiasedMHA


class GraphormerLayer(nn.Module):
    def __init__(self, feat_size, hidden_size, num_heads, attn_bias_type="add", norm_first=False, dropout=0.1,
                 attn_dropout=0.1, activation=nn.ReLU()):
        super().__init__()
        self.norm_first = norm_first
        self.attn = BiasedMHA(feat_size=feat_size, num_heads=num_heads, attn_bias_type=attn_bias_type, attn_drop=attn_dropout)
        self.ffn = nn.Sequential(nn.Linear(feat_size, hidden_size), activation, nn.Dropout(p=dropout),
                                 nn.Linear(hidden_size, feat_size), nn.Dropout(p=dropout))
        self.dropout = nn.Dropout(p=dropout)
        self.attn_layer_norm = nn.LayerNorm(feat_size)
        self.ffn_layer_norm = nn.LayerNorm(feat_size)

    def forward(self, nfeat, attn_bias=None, attn_mask=None):
        residual = nfeat
        if self.norm_first:
            nfeat = self.attn_layer_norm(nfeat)
        nfeat = self.attn(nfeat, attn_bias, attn_mask)
        nfeat = self.dropout(nfeat)
        nfeat = residual + nfeat
        if not self.norm_first:
            nfeat = self.attn_layer_norm(nfeat)
        residual = nfeat
        if self.norm_first:
            nfeat = self.ffn_layer_norm(nfeat)
        nfeat = self.ffn(nfeat)
        nfeat = residual + nfeat
        if not self.norm_first:
            nfeat = self.ffn_layer_norm(nfeat)
        return nfeat


class EGTLayer(nn.Module):
    """Edge-augmented graph transformer layer."""

    def __init__(self, feat_size, edge_feat_size, num_heads, num_virtual_nodes, dropout=0, attn_dropout=0,
                 activation=nn.ELU(), edge_update=True):
        super().__init__()
        self.num_heads, self.num_virtual_nodes, self.edge_update = num_heads, num_virtual_nodes, edge_update
        assert feat_size % num_heads == 0, "feat_size must be divisible by num_heads"
        self.dot_dim = feat_size // num_heads
        self.mha_ln_h = nn.LayerNorm(feat_size)
        self.mha_ln_e = nn.LayerNorm(edge_feat_size)
        self.edge_input = nn.Linear(edge_feat_size, num_heads)
        self.qkv_proj = nn.Linear(feat_size, feat_size * 3)
        self.gate = nn.Linear(edge_feat_size, num_heads)
        self.attn_dropout = nn.Dropout(attn_dropout)
        self.node_output = nn.Linear(feat_size, feat_size)
        self.mha_dropout_h = nn.Dropout(dropout)
        self.node_ffn = nn.Sequential(nn.LayerNorm(feat_size), nn.Linear(feat_size, feat_size), activation,
                                      nn.Linear(feat_size, feat_size), nn.Dropout(dropout))
        if self.edge_update:
            self.edge_output = nn.Linear(num_heads, edge_feat_size)
            self.mha_dropout_e = nn.Dropout(dropout)
            self.edge_ffn = nn.Sequential(nn.LayerNorm(edge_feat_size), nn.Linear(edge_feat_size, edge_feat_size),
                                          activation, nn.Linear(edge_feat_size, edge_feat_size), nn.Dropout(dropout))

    def forward(self, nfeat, efeat, mask=None):
        nfeat_r1, efeat_r1 = nfeat, efeat
        nfeat_ln = self.mha_ln_h(nfeat)
        efeat_ln = self.mha_ln_e(efeat)
        qkv = self.qkv_proj(nfeat_ln)
        e_bias = self.edge_input(efeat_ln)
        gates = self.gate(efeat_ln)
        bsz, N, _ = qkv.shape
        q_h, k_h, v_h = qkv.view(bsz, N, -1, self.num_heads).split(self.dot_dim, dim=2)
        attn_hat = torch.einsum("bldh,bmdh->blmh", q_h, k_h)
        attn_hat = attn_hat.clamp(-5, 5) + e_bias
        if mask is None:
            gates = torch.sigmoid(gates)
            attn_tild = F.softmax(attn_hat, dim=2) * gates
        else:
            gates = torch.sigmoid(gates + mask.unsqueeze(-1))
            attn_tild = F.softmax(attn_hat + mask.unsqueeze(-1), dim=2) * gates
        attn_tild = self.attn_dropout(attn_tild)
        v_attn = torch.einsum("blmh,bmkh->blkh", attn_tild, v_h)
        degrees = torch.sum(gates, dim=2, keepdim=True)
        degree_scalers = torch.log(1 + degrees)
        degree_scalers[:, : self.num_virtual_nodes] = 1.0
        v_attn = v_attn * degree_scalers
        v_attn = v_attn.reshape(bsz, N, self.num_heads * self.dot_dim)
        nfeat = self.node_output(v_attn)
        nfeat = self.mha_dropout_h(nfeat)
        nfeat = nfeat + nfeat_r1
        nfeat = nfeat + self.node_ffn(nfeat)
        if self.edge_update:
            efeat = self.edge_output(attn_hat)
            efeat = self.mha_dropout_e(efeat)
            efeat = efeat + efeat_r1
            efeat = efeat + self.edge_ffn(efeat)
            return nfeat, efeat
        return nfeat


class DegreeEncoder(nn.Module):
    def __init__(self, max_degree, embedding_dim, direction="both"):
        super().__init__()
        self.direction = direction
        if direction == "both":
            self.encoder1 = nn.Embedding(max_degree + 1, embedding_dim, padding_idx=0)
            self.encoder2 = nn.Embedding(max_degree + 1, embedding_dim, padding_idx=0)
        else:
            self.encoder = nn.Embedding(max_degree + 1, embedding_dim, padding_idx=0)
        self.max_degree = max_degree

    def forward(self, degrees):
        degrees = torch.clamp(degrees, min=0, max=self.max_degree)
        if self.direction == "in":
            return self.encoder(degrees)
        if self.direction == "out":
            return self.encoder(degrees)
        if self.direction == "both":
            return self.encoder1(degrees[0]) + self.encoder2(degrees[1])
        raise ValueError(f'Supported direction options: "in", "out" and "both", but got {self.direction}')


class PathEncoder(nn.Module):
    def __init__(self, max_len, feat_dim, num_heads=1):
        super().__init__()
        self.max_len, self.feat_dim, self.num_heads = max_len, feat_dim, num_heads
        self.embedding_table = nn.Embedding(max_len * num_heads, feat_dim)

    def forward(self, dist, path_data):
        shortest_distance = torch.clamp(dist, min=1, max=self.max_len)
        edge_embedding = self.embedding_table.weight.reshape(self.max_len, self.num_heads, -1)
        path_encoding = torch.div(torch.einsum("bxyld,lhd->bxyh", path_data, edge_embedding).permute(3, 0, 1, 2),
                                  shortest_distance).permute(1, 2, 3, 0)
        return path_encoding


class SpatialEncoder(nn.Module):
    def __init__(self, max_dist, num_heads=1):
        super().__init__()
        self.max_dist, self.num_heads = max_dist, num_heads
        self.embedding_table = nn.Embedding(max_dist + 2, num_heads, padding_idx=0)

    def forward(self, dist):
        return self.embedding_table(torch.clamp(dist, min=-1, max=self.max_dist) + 1)


class SpatialEncoder3d(nn.Module):
    def __init__(self, num_kernels, num_heads=1, max_node_type=100):
        super().__init__()
        self.num_kernels, self.num_heads, self.max_node_type = num_kernels, num_heads, max_node_type
        self.means = nn.Parameter(torch.empty(num_kernels))
        self.stds = nn.Parameter(torch.empty(num_kernels))
        self.linear_layer_1 = nn.Linear(num_kernels, num_kernels)
        self.linear_layer_2 = nn.Linear(num_kernels, num_heads)
        if max_node_type == 1:
            self.mul = nn.Embedding(1, 1, padding_idx=0)
            self.bias = nn.Embedding(1, 1, padding_idx=0)
        else:
            self.mul = nn.Embedding(max_node_type + 1, 2, padding_idx=0)
            self.bias = nn.Embedding(max_node_type + 1, 2, padding_idx=0)
        nn.init.uniform_(self.means, 0, 3)
        nn.init.uniform_(self.stds, 0, 3)
        nn.init.constant_(self.mul.weight, 0)
        nn.init.constant_(self.bias.weight, 1)

    def forward(self, coord, node_type=None):
        bsz, N = coord.shape[:2]
        euc_dist = torch.cdist(coord, coord, p=2.0)
        if node_type is None:
            node_type = euc_dist.new_zeros(bsz, N, N, 2).long() if self.max_node_type != 1 else euc_dist.new_zeros(bsz, N, N, 1).long()
            mul = self.mul.weight[0, 0] if self.max_node_type == 1 else self.mul(node_type).sum(-1)
            bias = self.bias.weight[0, 0] if self.max_node_type == 1 else self.bias(node_type).sum(-1)
        else:
            src_node_type = node_type.unsqueeze(-1).repeat(1, 1, N)
            tgt_node_type = node_type.unsqueeze(1).repeat(1, N, 1)
            mul_edge = self.mul(torch.stack([src_node_type, tgt_node_type], -1))
            bias_edge = self.bias(torch.stack([src_node_type, tgt_node_type], -1))
            mul = mul_edge.sum(-1).sum(-1) if mul_edge.dim() == 5 else mul_edge.sum(-1)
            bias = bias_edge.sum(-1).sum(-1) if bias_edge.dim() == 5 else bias_edge.sum(-1)
        scaled_dist = (mul * euc_dist + bias).unsqueeze(-1)
        sqrt_2pi = math.sqrt(2 * math.pi)
        gaussian_kernel = (scaled_dist - self.means) / self.stds.abs().clamp(min=1e-2)
        gaussian_kernel = torch.exp(-0.5 * gaussian_kernel ** 2) / (sqrt_2pi * self.stds.abs().clamp(min=1e-2))
        encoding = self.linear_layer_2(F.gelu(self.linear_layer_1(gaussian_kernel)))
        return encoding


class LapPosEncoder(nn.Module):
    def __init__(self, model_type, num_layer, k, dim, n_head=1, batch_norm=False, num_post_layer=0):
        super().__init__()
        self.model_type, self.k = model_type, k
        if model_type == "Transformer":
            self.linear = nn.Linear(2, dim)
            encoder_layer = nn.TransformerEncoderLayer(d_model=dim, nhead=n_head, batch_first=True)
            self.pe_encoder = nn.TransformerEncoder(encoder_layer, num_layers=num_layer)
        elif model_type == "DeepSet":
            layers = []
            if num_layer == 1:
                layers.append(nn.ReLU())
            else:
                self.linear = nn.Linear(2, 2 * dim)
                layers.append(nn.ReLU())
                for _ in range(num_layer - 2):
                    layers.append(nn.Linear(2 * dim, 2 * dim))
                    layers.append(nn.ReLU())
                layers.append(nn.Linear(2 * dim, dim))
                layers.append(nn.ReLU())
            self.pe_encoder = nn.Sequential(*layers)
        else:
            raise ValueError(f"model_type should be one of ['Transformer', 'DeepSet'], but got {model_type}")
        self.raw_norm = nn.BatchNorm1d(k) if batch_norm else None
        if num_post_layer > 0:
            layers = []
            if num_post_layer == 1:
                layers += [nn.Linear(dim, dim), nn.ReLU()]
            else:
                layers += [nn.Linear(dim, 2 * dim), nn.ReLU()]
                for _ in range(num_post_layer - 2):
                    layers += [nn.Linear(2 * dim, 2 * dim), nn.ReLU()]
                layers += [nn.Linear(2 * dim, dim), nn.ReLU()]
            self.post_mlp = nn.Sequential(*layers)
        else:
            self.post_mlp = None

    def forward(self, eigvals, eigvecs):
        pos_encoding = torch.cat((eigvecs.unsqueeze(2), eigvals.unsqueeze(2)), dim=2).float()
        empty_mask = torch.isnan(pos_encoding)
        pos_encoding[empty_mask] = 0
        if self.raw_norm:
            pos_encoding = self.raw_norm(pos_encoding)
        pos_encoding = self.linear(pos_encoding)
        if self.model_type == "Transformer":
            pos_encoding = self.pe_encoder(src=pos_encoding, src_key_padding_mask=empty_mask[:, :, 1])
        else:
            pos_encoding = self.pe_encoder(pos_encoding)
        pos_encoding[empty_mask[:, :, 1]] = 0
        pos_encoding = torch.sum(pos_encoding, 1, keepdim=False)
        if self.post_mlp:
            pos_encoding = self.post_mlp(pos_encoding)
        return pos_encoding
