"""Graph-level readout modules: Sum / Avg / Max / Sort pooling, WeightAndSum, GlobalAttentionPooling, Set2Set and the
Set Transformer encoder / decoder.

API parity: /root/reference/python/dgl/nn/pytorch/glob.py (class names, constructor arguments, parameter names). The
reference stores features in node frames and calls readout functions on them, and its set attention pads every set to
the longest one and masks; here

* every readout is a SEGMENT operation on the packed node tensor (``ops.segment_reduce`` / ``ops.segment_softmax`` over
  ``batch_num_nodes``) - no frames, no ``local_scope``;
* set attention (``MultiHeadAttention``) is computed on the packed representation too: the (query, key) pairs of all
  sets are enumerated once, scores are a gathered dot product, the softmax is a segment softmax over each query's keys
  and the weighted sum a segment reduction - no padding to the longest set, no masks, work proportional to
  sum_i |X_i| |M_i|.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn as nn

from .. import ops

__all__ = ["SumPooling", "AvgPooling", "MaxPooling", "SortPooling", "GlobalAttentionPooling", "Set2Set",
           "SetTransformerEncoder", "SetTransformerDecoder", "WeightAndSum"]


def _sizes(graph, ntype=None):
    return graph.batch_num_nodes(ntype) if ntype is not None else graph.batch_num_nodes()


def _pool(graph, feat, how, ntype=None):
    return ops.segment_reduce(_sizes(graph, ntype).to(feat.device), feat, how)


def _spread(graph, per_graph):
    """Repeat one row per graph to one row per node."""
    sizes = _sizes(graph).to(per_graph.device)
    return torch.repeat_interleave(per_graph, sizes, dim=0, output_size=graph.num_nodes())


class SumPooling(nn.Module):
    """``r_g = sum_{v in g} x_v``"""

    def forward(self, graph, feat, ntype=None):
        return _pool(graph, feat, "sum", ntype)


class AvgPooling(nn.Module):
    """``r_g = mean_{v in g} x_v``"""

    def forward(self, graph, feat, ntype=None):
        return _pool(graph, feat, "mean", ntype)


class MaxPooling(nn.Module):
    """``r_g = max_{v in g} x_v`` (element-wise)"""

    def forward(self, graph, feat, ntype=None):
        return _pool(graph, feat, "max", ntype)


class SortPooling(nn.Module):
    """Sort every node's features ascending, keep the ``k`` nodes with the largest last channel per graph (zero padded)
    and concatenate them: [batch, k * d]."""

    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, graph, feat):
        from ..readout import topk_nodes
        ordered = feat.sort(dim=-1)[0]
        with graph.local_scope():
            graph.ndata["_sortpool"] = ordered
            top = topk_nodes(graph, "_sortpool", self.k, sortby=-1)[0]
        return top.reshape(-1, self.k * feat.shape[-1])


class GlobalAttentionPooling(nn.Module):
    """``r_g = sum_v softmax_g(gate(x_v)) * feat_nn(x_v)``"""

    def __init__(self, gate_nn, feat_nn=None):
        super().__init__()
        self.gate_nn, self.feat_nn = gate_nn, feat_nn

    def forward(self, graph, feat, get_attention=False):
        score = self.gate_nn(feat)
        if score.shape[-1] != 1:
            raise AssertionError("The output of gate_nn should have size 1 at the last axis.")
        sizes = _sizes(graph).to(feat.device)
        weight = ops.segment_softmax(sizes, score)
        values = feat if self.feat_nn is None else self.feat_nn(feat)
        pooled = ops.segment_reduce(sizes, values * weight, "sum")
        return (pooled, weight) if get_attention else pooled


class WeightAndSum(nn.Module):
    """``r_g = sum_v sigmoid(w . x_v + b) x_v``"""

    def __init__(self, in_feats):
        super().__init__()
        self.in_feats = in_feats
        self.atom_weighting = nn.Sequential(nn.Linear(in_feats, 1), nn.Sigmoid())

    def forward(self, g, feats):
        return _pool(g, feats * self.atom_weighting(feats), "sum")


class Set2Set(nn.Module):
    """Order-invariant readout: an LSTM emits a query per graph, attention over the graph's nodes answers it."""

    def __init__(self, input_dim, n_iters, n_layers):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, 2 * input_dim
        self.n_iters, self.n_layers = n_iters, n_layers
        self.lstm = nn.LSTM(self.output_dim, input_dim, n_layers)
        self.reset_parameters()

    def reset_parameters(self):
        self.lstm.reset_parameters()

    def forward(self, graph, feat):
        B = graph.batch_size
        sizes = _sizes(graph).to(feat.device)
        state = tuple(feat.new_zeros(self.n_layers, B, self.input_dim) for _ in range(2))
        q_star = feat.new_zeros(B, self.output_dim)
        for _ in range(self.n_iters):
            q, state = self.lstm(q_star.unsqueeze(0), state)
            q = q.squeeze(0)
            match = (feat * _spread(graph, q)).sum(-1, keepdim=True)
            read = ops.segment_reduce(sizes, feat * ops.segment_softmax(sizes, match), "sum")
            q_star = torch.cat([q, read], dim=-1)
        return q_star

    def extra_repr(self):
        return f"n_iters={self.n_iters}"


# ---------------------------------------------------------------------------------------------------------------------
#  Set Transformer on packed sets
# ---------------------------------------------------------------------------------------------------------------------
def _cross_pairs(len_x, len_m, device):
    """All (query, key) index pairs between set i of X and set i of M, grouped by query.
    Returns (key index per pair, keys per query [Nx])."""
    lx = torch.as_tensor(len_x, dtype=torch.long, device=device)
    lm = torch.as_tensor(len_m, dtype=torch.long, device=device)
    n_x = int(sum(len_x))
    n_pairs = int(sum(a * b for a, b in zip(len_x, len_m)))
    set_of_q = torch.repeat_interleave(torch.arange(lx.shape[0], device=device), lx, output_size=n_x)
    keys_per_q = lm[set_of_q]
    first_key = (torch.cumsum(lm, 0) - lm)[set_of_q]                              # offset of the query's set in M
    q_of_pair = torch.repeat_interleave(torch.arange(n_x, device=device), keys_per_q, output_size=n_pairs)
    rank = torch.arange(n_pairs, device=device) - (torch.cumsum(keys_per_q, 0) - keys_per_q)[q_of_pair]
    return q_of_pair, first_key[q_of_pair] + rank, keys_per_q


class MultiHeadAttention(nn.Module):
    """Multi-head attention block (attention + residual LayerNorm + feed-forward) between packed sets."""

    def __init__(self, d_model, num_heads, d_head, d_ff, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.d_model, self.num_heads, self.d_head, self.d_ff = d_model, num_heads, d_head, d_ff
        inner = num_heads * d_head
        self.proj_q = nn.Linear(d_model, inner, bias=False)
        self.proj_k = nn.Linear(d_model, inner, bias=False)
        self.proj_v = nn.Linear(d_model, inner, bias=False)
        self.proj_o = nn.Linear(inner, d_model, bias=False)
        self.ffn = nn.Sequential(nn.Linear(d_model, d_ff), nn.ReLU(), nn.Dropout(dropouth), nn.Linear(d_ff, d_model))
        self.droph, self.dropa = nn.Dropout(dropouth), nn.Dropout(dropouta)   # kept for state / repr parity; unused
        self.norm_in, self.norm_inter = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.reset_parameters()

    def reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, x, mem, lengths_x, lengths_mem):
        H, d = self.num_heads, self.d_head
        q_idx, k_idx, keys_per_q = _cross_pairs(list(lengths_x), list(lengths_mem), x.device)
        Q = self.proj_q(x).view(-1, H, d)
        K = self.proj_k(mem).view(-1, H, d)
        V = self.proj_v(mem).view(-1, H, d)
        score = (Q[q_idx] * K[k_idx]).sum(-1) / math.sqrt(d)                      # [pairs, H]
        alpha = ops.segment_softmax(keys_per_q, score)
        mixed = ops.segment_reduce(keys_per_q, alpha.unsqueeze(-1) * V[k_idx], "sum")   # [Nx, H, d]
        x = self.norm_in(x + self.proj_o(mixed.reshape(-1, H * d)))
        return self.norm_inter(x + self.ffn(x))


class SetAttentionBlock(nn.Module):
    """SAB: self-attention inside every set."""

    def __init__(self, d_model, num_heads, d_head, d_ff, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.mha = MultiHeadAttention(d_model, num_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)

    def forward(self, feat, lengths):
        return self.mha(feat, feat, lengths, lengths)


def _warn_single(name, what):
    warnings.warn(f"if {name} is set to 1, the parameters corresponding to query and key projections would not get "
                  f"updated during training ({what}).")


class InducedSetAttentionBlock(nn.Module):
    """ISAB: ``m`` learned inducing points attend to the set, then the set attends to their summaries (O(n m))."""

    def __init__(self, m, d_model, num_heads, d_head, d_ff, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.m, self.d_model = m, d_model
        if m == 1:
            _warn_single("m", "one inducing point")
        self.inducing_points = nn.Parameter(torch.empty(m, d_model))
        self.mha = nn.ModuleList([MultiHeadAttention(d_model, num_heads, d_head, d_ff, dropouth=dropouth,
                                                     dropouta=dropouta) for _ in range(2)])
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.inducing_points)

    def forward(self, feat, lengths):
        n_sets = len(lengths)
        per_set = [self.m] * n_sets
        summary = self.mha[0](self.inducing_points.repeat(n_sets, 1), feat, per_set, lengths)
        return self.mha[1](feat, summary, lengths, per_set)


class PMALayer(nn.Module):
    """Pooling by multi-head attention: ``k`` learned seed vectors query the (feed-forwarded) set."""

    def __init__(self, k, d_model, num_heads, d_head, d_ff, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.k, self.d_model = k, d_model
        if k == 1:
            _warn_single("k", "one seed vector")
        self.seed_vectors = nn.Parameter(torch.empty(k, d_model))
        self.mha = MultiHeadAttention(d_model, num_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)
        self.ffn = nn.Sequential(nn.Linear(d_model, d_ff), nn.ReLU(), nn.Dropout(dropouth), nn.Linear(d_ff, d_model))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.seed_vectors)

    def forward(self, feat, lengths):
        n_sets = len(lengths)
        return self.mha(self.seed_vectors.repeat(n_sets, 1), self.ffn(feat), [self.k] * n_sets, lengths)


_ENCODER_BLOCKS = {
    "sab": lambda m, *a, **kw: SetAttentionBlock(*a, **kw),
    "isab": lambda m, *a, **kw: InducedSetAttentionBlock(m, *a, **kw),
}


class SetTransformerEncoder(nn.Module):
    """``n_layers`` SAB or ISAB blocks over the node sets of a batched graph."""

    def __init__(self, d_model, n_heads, d_head, d_ff, n_layers=1, block_type="sab", m=None, dropouth=0.0,
                 dropouta=0.0):
        super().__init__()
        if block_type not in _ENCODER_BLOCKS:
            raise KeyError(f"Unrecognized block type {block_type}: we only support sab/isab")
        if block_type == "isab" and m is None:
            raise KeyError("The number of inducing points is not specified in ISAB block.")
        self.n_layers, self.block_type, self.m = n_layers, block_type, m
        make = _ENCODER_BLOCKS[block_type]
        self.layers = nn.ModuleList([make(m, d_model, n_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)
                                     for _ in range(n_layers)])

    def forward(self, graph, feat):
        lengths = graph.batch_num_nodes().tolist()
        for block in self.layers:
            feat = block(feat, lengths)
        return feat


class SetTransformerDecoder(nn.Module):
    """PMA pooling to ``k`` vectors per graph followed by ``n_layers`` SABs; returns [batch, k * d_model]."""

    def __init__(self, d_model, num_heads, d_head, d_ff, n_layers, k, dropouth=0.0, dropouta=0.0):
        super().__init__()
        self.n_layers, self.k, self.d_model = n_layers, k, d_model
        self.pma = PMALayer(k, d_model, num_heads, d_head, d_ff, dropouth=dropouth, dropouta=dropouta)
        self.layers = nn.ModuleList([SetAttentionBlock(d_model, num_heads, d_head, d_ff, dropouth=dropouth,
                                                       dropouta=dropouta) for _ in range(n_layers)])

    def forward(self, graph, feat):
        pooled = self.pma(feat, graph.batch_num_nodes().tolist())
        per_set = [self.k] * graph.batch_size
        for block in self.layers:
            pooled = block(pooled, per_set)
        return pooled.view(graph.batch_size, self.k * self.d_model)
