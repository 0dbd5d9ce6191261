"""Global pooling modules (parity: /root/reference/python/dgl/nn/pytorch/glob.py): SumPooling, AvgPooling,
MaxPooling, SortPooling, WeightAndSum, GlobalAttentionPooling, Set2Set, SetTransformerEncoder/Decoder.
Readouts use the segment-reduce kernel through dgl_b200.readout."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..readout import (broadcast_nodes, max_nodes, mean_nodes, softmax_nodes, sum_nodes, topk_nodes)

__all__ = ["SumPooling", "AvgPooling", "MaxPooling", "SortPooling", "GlobalAttentionPooling", "Set2Set",
           "SetTransformerEncoder", "SetTransformerDecoder", "WeightAndSum"]


class SumPooling(nn.Module):
    def forward(self, graph, feat, ntype=None):
        with graph.local_scope():
            graph.nodes[ntype].data["h"] = feat
            return sum_nodes(graph, "h", ntype=ntype)


class AvgPooling(nn.Module):
    def forward(self, graph, feat, ntype=None):
        with graph.local_scope():
            graph.nodes[ntype].data["h"] = feat
            return mean_nodes(graph, "h", ntype=ntype)


class MaxPooling(nn.Module):
    def forward(self, graph, feat, ntype=None):
        with graph.local_scope():
            graph.nodes[ntype].data["h"] = feat
            return max_nodes(graph, "h", ntype=ntype)


class SortPooling(nn.Module):
    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, graph, feat):
        with graph.local_scope():
            feat, _ = feat.sort(dim=-1)
            graph.ndata["h"] = feat
            ret = topk_nodes(graph, "h", self.k, sortby=-1)[0].view(-1, self.k * feat.shape[-1])
            return ret


class GlobalAttentionPooling(nn.Module):
    def __init__(self, gate_nn, feat_nn=None):
        super().__init__()
        self.gate_nn = gate_nn
        self.feat_nn = feat_nn

    def forward(self, graph, feat, get_attention=False):
        with graph.local_scope():
            gate = self.gate_nn(feat)
            assert gate.shape[-1] == 1, "The output of gate_nn should have size 1 at the last axis."
            feat = self.feat_nn(feat) if self.feat_nn else feat
            graph.ndata["gate"] = gate
            gate = softmax_nodes(graph, "gate")
            graph.ndata.pop("gate")
            graph.ndata["r"] = feat * gate
            readout = sum_nodes(graph, "r")
            graph.ndata.pop("r")
            if get_attention:
                return readout, gate
            return readout


class WeightAndSum(nn.Module):
    def __init__(self, in_feats):
        super().__init__()
        self.in_feats = in_feats
        self.atom_weighting = nn.Sequential(nn.Linear(in_feats, 1), nn.Sigmoid())

    def forward(self, g, feats):
        with g.local_scope():
            g.ndata["h"] = feats
            g.ndata["w"] = self.atom_weighting(g.ndata["h"])
            return sum_nodes(g, "h", "w")


class Set2Set(nn.Module):
    def __init__(self, input_dim, n_iters, n_layers):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = 2 * input_dim
        self.n_iters = n_iters
        self.n_layers = n_layers
        self.lstm = nn.LSTM(self.output_dim, self.input_dim, n_layers)
        self.reset_parameters()

    def reset_parameters(self):
        self.lstm.reset_parameters()

    def forward(self, graph, feat):
        with graph.local_scope():
            batch_size = graph.batch_size
            h = (feat.new_zeros((self.n_layers, batch_size, self.input_dim)),
                 feat.new_zeros((self.n_layers, batch_size, self.input_dim)))
            q_star = feat.new_zeros(batch_size, self.output_dim)
            for _ in range(self.n_iters):
                q, h = self.lstm(q_star.unsqueeze(0), h)
                q = q.view(batch_size, self.input_dim)
                e = (feat * broadcast_nodes(graph, q)).sum(dim=-1, keepdim=True)
                graph.ndata["e"] = e
                alpha = softmax_nodes(graph, "e")
                graph.ndata["r"] = feat * alpha
                readout = sum_nodes(graph, "r")
                q_star = torch.cat([q, readout], dim=-1)
            return q_star

    def extra_repr(self):
        return "n_iters={n_iters}".format(**self.__dict__)


def _gen_mask(lengths_x, lengths_y, max_len_x, max_len_y):
    device = lengths_x.device
    x_mask = torch.arange(max_len_x, device=device).unsqueeze(0) < lengths_x.unsqueeze(1)
    y_mask = torch.arange(max_len_y, device=device).unsqueeze(0) < lengths_y.unsqueeze(1)
    return (x_mask.unsqueeze(-1) & y_mask.unsqueeze(-2)).unsqueeze(1)


def _pad_packed(x, lengths):
    B = lengths.shape[0]
    maxlen = int(lengths.max()) if B else 0
    out = x.new_zeros((B, maxlen) + tuple(x.shape[1:]))
    idx_b = torch.repeat_interleave(torch.arange(B, device=x.device), lengths)
    pos = torch.arange(x.shape[0], device=x.device) - torch.repeat_interleave(torch.cumsum(lengths, 0) - lengths, lengths)
    out[idx_b, pos] = x
    return out, idx_b, pos


class MultiHeadAttention(nn.Module):
    def __init__(self, d_model, num_heads, d_head, d_ff, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.d_model, self.num_heads, self.d_head, self.d_ff = d_model, num_heads, d_head, d_ff
        self.proj_q = nn.Linear(d_model, num_heads * d_head, bias=False)
        self.proj_k = nn.Linear(d_model, num_heads * d_head, bias=False)
        self.proj_v = nn.Linear(d_model, num_heads * d_head, bias=False)
        self.proj_o = nn.Linear(num_heads * d_head, d_model, bias=False)
        self.ffn = nn.Sequential(nn.Linear(d_model, d_ff), nn.ReLU(), nn.Dropout(dropouth), nn.Linear(d_ff, d_model))
        self.droph = nn.Dropout(dropouth)
        self.dropa = nn.Dropout(dropouta)
        self.norm_in = nn.LayerNorm(d_model)
        self.norm_inter = nn.LayerNorm(d_model)
        self.reset_parameters()

    def reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, x, mem, lengths_x, lengths_mem):
        batch_size = len(lengths_x)
        lx = torch.as_tensor(lengths_x, device=x.device).long()
        lm = torch.as_tensor(lengths_mem, device=x.device).long()
        max_len_x, max_len_mem = int(lx.max()), int(lm.max())
        queries = self.proj_q(x).view(-1, self.num_heads, self.d_head)
        keys = self.proj_k(mem).view(-1, self.num_heads, self.d_head)
        values = self.proj_v(mem).view(-1, self.num_heads, self.d_head)
        queries, bx, px = _pad_packed(queries, lx)
        keys, _, _ = _pad_packed(keys, lm)
        values, _, _ = _pad_packed(values, lm)
        e = torch.einsum("bxhd,byhd->bhxy", queries, keys) / np.sqrt(self.d_head)
        mask = _gen_mask(lx, lm, max_len_x, max_len_mem)
        e = e.masked_fill(mask == 0, -float("inf"))
        alpha = torch.softmax(e, dim=-1)
        alpha = alpha.masked_fill(mask == 0, 0.0)
        out = torch.einsum("bhxy,byhd->bxhd", alpha, values)
        out = self.proj_o(out.contiguous().view(batch_size, max_len_x, self.num_heads * self.d_head))
        out = out[bx, px]
        x = self.norm_in(x + out)
        x = self.norm_inter(x + self.ffn(x))
        return x


class SetAttentionBlock(nn.Module):
    def __init__(self, d_model, num_heads, d_head, d_ff, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.mha = MultiHeadAttention(d_model, num_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)

    def forward(self, feat, lengths):
        return self.mha(feat, feat, lengths, lengths)


class InducedSetAttentionBlock(nn.Module):
    def __init__(self, m, d_model, num_heads, d_head, d_ff, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.m = m
        if m == 1:
            import warnings
            warnings.warn("if m is set to 1, the parameters corresponding to query and key projections would not get updated during training.")
        self.d_model = d_model
        self.inducing_points = nn.Parameter(torch.FloatTensor(m, d_model))
        self.mha = nn.ModuleList([MultiHeadAttention(d_model, num_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)
                                  for _ in range(2)])
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.inducing_points)

    def forward(self, feat, lengths):
        batch_size = len(lengths)
        query = self.inducing_points.repeat(batch_size, 1)
        memory = self.mha[0](query, feat, [self.m] * batch_size, lengths)
        return self.mha[1](feat, memory, lengths, [self.m] * batch_size)


class PMALayer(nn.Module):
    def __init__(self, k, d_model, num_heads, d_head, d_ff, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.k = k
        if k == 1:
            import warnings
            warnings.warn("if k is set to 1, the parameters corresponding to query and key projections would not get updated during training.")
        self.d_model = d_model
        self.seed_vectors = nn.Parameter(torch.FloatTensor(k, d_model))
        self.mha = MultiHeadAttention(d_model, num_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)
        self.ffn = nn.Sequential(nn.Linear(d_model, d_ff), nn.ReLU(), nn.Dropout(dropouth), nn.Linear(d_ff, d_model))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.seed_vectors)

    def forward(self, feat, lengths):
        batch_size = len(lengths)
        query = self.seed_vectors.repeat(batch_size, 1)
        return self.mha(query, self.ffn(feat), [self.k] * batch_size, lengths)


class SetTransformerEncoder(nn.Module):
    def __init__(self, d_model, n_heads, d_head, d_ff, n_layers=1, block_type="sab", m=None, dropouth=0.0,
                 dropouta=0.0):
        super().__init__()
        self.n_layers, self.block_type, self.m = n_layers, block_type, m
        layers = []
        if block_type == "isab" and m is None:
            raise KeyError("The number of inducing points is not specified in ISAB block.")
        for _ in range(n_layers):
            if block_type == "sab":
                layers.append(SetAttentionBlock(d_model, n_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta))
            elif block_type == "isab":
                layers.append(InducedSetAttentionBlock(m, d_model, n_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta))
            else:
                raise KeyError("Unrecognized block type {}: we only support sab/isab".format(block_type))
        self.layers = nn.ModuleList(layers)

    def forward(self, graph, feat):
        lengths = graph.batch_num_nodes().tolist()
        for layer in self.layers:
            feat = layer(feat, lengths)
        return feat


class SetTransformerDecoder(nn.Module):
    def __init__(self, d_model, num_heads, d_head, d_ff, n_layers, k, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.n_layers, self.k, self.d_model = n_layers, k, d_model
        self.pma = PMALayer(k, d_model, num_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)
        self.layers = nn.ModuleList([SetAttentionBlock(d_model, num_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)
                                     for _ in range(n_layers)])

    def forward(self, graph, feat):
        len_pma = graph.batch_num_nodes().tolist()
        len_sab = [self.k] * graph.batch_size
        feat = self.pma(feat, len_pma)
        for layer in self.layers:
            feat = layer(feat, len_sab)
        return feat.view(graph.batch_size, self.k * self.d_model)
