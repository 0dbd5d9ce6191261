"""R-GCN (Schlichtkrull et al.) for node classification on a heterograph flattened to typed edges
(reference model: examples/pytorch/rgcn/entity.py - RelGraphConv with basis regularisation)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn import RelGraphConv


class RGCN(nn.Module):
    def __init__(self, num_nodes, hid_size, out_size, num_rels, num_bases=None, in_size=None, num_layers=2,
                 regularizer="basis", dropout=0.0):
        super().__init__()
        # featureless nodes get a learnable embedding (entity.py:35-48)
        self.emb = nn.Embedding(num_nodes, hid_size) if in_size is None else None
        first = hid_size if in_size is None else in_size
        dims = [first] + [hid_size] * (num_layers - 1) + [out_size]
        self.layers = nn.ModuleList(
            RelGraphConv(dims[i], dims[i + 1], num_rels, regularizer=regularizer, num_bases=num_bases, self_loop=True,
                         dropout=dropout if i < num_layers - 1 else 0.0) for i in range(num_layers))

    def forward(self, g, etypes, x=None, nids=None):
        """``g`` homogeneous graph (or list of blocks) whose edges carry relation ids ``etypes``."""
        if self.emb is not None:
            h = self.emb.weight if nids is None else self.emb(nids)
        else:
            h = x
        for i, layer in enumerate(self.layers):
            gi = g[i] if isinstance(g, (list, tuple)) else g
            ei = etypes[i] if isinstance(etypes, (list, tuple)) else etypes
            h = layer(gi, h, ei)
            if i < len(self.layers) - 1:
                h = F.relu(h)
        return h
