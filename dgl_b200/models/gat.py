"""GAT (Velickovic et al.) - multi-head attention via the fused edge-softmax + g-SDDMM/g-SpMM kernels
(reference model: examples/pytorch/gat/train.py)."""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F

from ..nn import GATConv


class GAT(nn.Module):
    def __init__(self, in_size, hid_size, out_size, heads=(8, 1), feat_drop=0.6, attn_drop=0.6):
        super().__init__()
        self.layers = nn.ModuleList()
        last = in_size
        for i, h in enumerate(heads):
            final = i == len(heads) - 1
            self.layers.append(GATConv(last, out_size if final else hid_size, h, feat_drop=feat_drop, attn_drop=attn_drop,
                                       activation=None if final else F.elu, allow_zero_in_degree=True))
            last = hid_size * h

    def forward(self, g, x):
        h = x
        for i, layer in enumerate(self.layers):
            h = layer(g[i] if isinstance(g, (list, tuple)) else g, h)
            h = h.mean(1) if i == len(self.layers) - 1 else h.flatten(1)
        return h
