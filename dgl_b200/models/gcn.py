"""GCN (Kipf & Welling) - the reference's examples/pytorch/gcn/train.py model on top of nn.GraphConv."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn import GraphConv


class GCN(nn.Module):
    def __init__(self, in_size, hid_size, out_size, num_layers=2, dropout=0.5):
        super().__init__()
        dims = [in_size] + [hid_size] * (num_layers - 1) + [out_size]
        self.layers = nn.ModuleList(
            GraphConv(dims[i], dims[i + 1], activation=F.relu if i < num_layers - 1 else None, allow_zero_in_degree=True)
            for i in range(num_layers))
        self.dropout = nn.Dropout(dropout)

    def forward(self, g, features):
        """``g`` is a graph (full-graph training) or a list of blocks (mini-batch)."""
        h = features
        for i, layer in enumerate(self.layers):
            if i != 0:
                h = self.dropout(h)
            h = layer(g[i] if isinstance(g, (list, tuple)) else g, h)
        return h
