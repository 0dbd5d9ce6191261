"""SpotTarget (Guo et al., "Rethinking Target Edge Leakage"): during link-prediction training only those target edges are
hidden from message passing that touch a LOW-degree endpoint - for high-degree endpoints removing one edge changes little
and keeping it avoids a train / test distribution shift.

Callable for the ``exclude=`` argument of ``as_edge_prediction_sampler`` (reference
python/dgl/dataloading/spot_target.py:26-90)."""
from __future__ import annotations

import torch


class SpotTarget:
    def __init__(self, g, exclude, degree_threshold=10, reverse_eids=None, reverse_etypes=None):
        self.g, self.exclude, self.degree_threshold = g, exclude, degree_threshold
        self.reverse_eids, self.reverse_etypes = reverse_eids, reverse_etypes

    def __call__(self, seed_edges):
        u, v = self.g.find_edges(seed_edges)
        weakest = torch.minimum(self.g.in_degrees(u), self.g.in_degrees(v))     # the sparser endpoint decides
        hidden = seed_edges[weakest < self.degree_threshold]
        if self.exclude == "reverse_id" and self.reverse_eids is not None:
            partner = self.reverse_eids.to(hidden.device)[hidden.long()]
            hidden = torch.cat([hidden, partner])
        return hidden
