"""Typed / heterogeneous linear layers (parity: /root/reference/python/dgl/nn/pytorch/linear.py:13-212,
hetero.py HeteroLinear/HeteroEmbedding). TypedLinear picks segment_mm when inputs are sorted by type — with a
device-side seglen (no ``.cpu()`` sync as in the reference linear.py:205-207) — else gather_mm."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..ops import gather_mm, segment_mm


class TypedLinear(nn.Module):
    r"""y_i = W_{t_i} x_i with optional basis / block-diagonal weight regularisation."""

    def __init__(self, in_size, out_size, num_types, regularizer=None, num_bases=None):
        super().__init__()
        self.in_size, self.out_size, self.num_types = in_size, out_size, num_types
        if regularizer is None:
            self.W = nn.Parameter(torch.Tensor(num_types, in_size, out_size))
        elif regularizer == "basis":
            if num_bases is None:
                raise ValueError('Missing "num_bases" for basis regularization.')
            self.W = nn.Parameter(torch.Tensor(num_bases, in_size, out_size))
            self.coeff = nn.Parameter(torch.Tensor(num_types, num_bases))
            self.num_bases = num_bases
        elif regularizer == "bdd":
            if num_bases is None:
                raise ValueError('Missing "num_bases" for bdd regularization.')
            if in_size % num_bases != 0 or out_size % num_bases != 0:
                raise ValueError("Input and output sizes must be divisible by num_bases.")
            self.submat_in = in_size // num_bases
            self.submat_out = out_size // num_bases
            self.W = nn.Parameter(torch.Tensor(num_types, num_bases * self.submat_in * self.submat_out))
            self.num_bases = num_bases
        else:
            raise ValueError(f'Supported regularizer options: "basis", "bdd", but got {regularizer}')
        self.regularizer = regularizer
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            if self.regularizer is None:
                nn.init.uniform_(self.W, -1 / math.sqrt(self.in_size), 1 / math.sqrt(self.in_size))
            elif self.regularizer == "basis":
                nn.init.uniform_(self.W, -1 / math.sqrt(self.in_size), 1 / math.sqrt(self.in_size))
                nn.init.xavier_uniform_(self.coeff, gain=nn.init.calculate_gain("relu"))
            elif self.regularizer == "bdd":
                nn.init.uniform_(self.W, -1 / math.sqrt(self.submat_in), 1 / math.sqrt(self.submat_in))

    def get_weight(self):
        if self.regularizer is None:
            return self.W
        if self.regularizer == "basis":
            W = self.W.view(self.num_bases, self.in_size * self.out_size)
            return (self.coeff @ W).view(self.num_types, self.in_size, self.out_size)
        if self.regularizer == "bdd":
            return self.W
        raise ValueError(self.regularizer)

    def forward(self, x, x_type, sorted_by_type=False):
        w = self.get_weight()
        if self.regularizer == "bdd":
            w = w.index_select(0, x_type.long()).view(-1, self.submat_in, self.submat_out)
            x = x.view(-1, 1, self.submat_in)
            return torch.bmm(x, w).view(-1, self.out_size)
        if sorted_by_type:
            seglen = torch.bincount(x_type.long(), minlength=self.num_types)
            return segment_mm(x, w, seglen_a=seglen)
        return gather_mm(x, w, idx_b=x_type)

    def __repr__(self):
        if self.regularizer is None:
            return f"TypedLinear(in_size={self.in_size}, out_size={self.out_size}, num_types={self.num_types})"
        return (f"TypedLinear(in_size={self.in_size}, out_size={self.out_size}, num_types={self.num_types}, "
                f"regularizer={self.regularizer}, num_bases={self.num_bases})")


class HeteroLinear(nn.Module):
    """One linear layer per key of a feature dict."""

    def __init__(self, in_size, out_size, bias=True):
        super().__init__()
        self.linears = nn.ModuleDict()
        for typ, typ_in_size in in_size.items():
            self.linears[str(typ)] = nn.Linear(typ_in_size, out_size, bias=bias)

    def forward(self, feat):
        return {typ: self.linears[str(typ)](typ_feat) for typ, typ_feat in feat.items()}


class HeteroEmbedding(nn.Module):
    """One embedding table per key."""

    def __init__(self, num_embeddings, embedding_dim):
        super().__init__()
        self.embeds = nn.ModuleDict()
        self.raw_keys = {}
        for typ, typ_num_rows in num_embeddings.items():
            self.embeds[str(typ)] = nn.Embedding(typ_num_rows, embedding_dim)
            self.raw_keys[str(typ)] = typ

    @property
    def weight(self):
        return {self.raw_keys[typ]: emb.weight for typ, emb in self.embeds.items()}

    def reset_parameters(self):
        for typ in self.embeds.keys():
            nn.init.xavier_uniform_(self.embeds[typ].weight)

    def forward(self, input_ids):
        return {typ: self.embeds[str(typ)](ids) for typ, ids in input_ids.items()}
