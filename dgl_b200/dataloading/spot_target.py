"""SpotTarget exclusion callable (parity: /root/reference/python/dgl/dataloading/spot_target.py):
exclude the target edges (and their reverses) only when incident to a low-degree node."""
from __future__ import annotations

import torch


class SpotTarget:
    def __init__(self, g, exclude, degree_threshold=10, reverse_eids=None, reverse_etypes=None):
        self.g = g
        self.exclude = exclude
        self.degree_threshold = degree_threshold
        self.reverse_eids = reverse_eids
        self.reverse_etypes = reverse_etypes

    def __call__(self, seed_edges):
        g = self.g
        src, dst = g.find_edges(seed_edges)
        head_degree = g.in_degrees(src)
        tail_degree = g.in_degrees(dst)
        degree = torch.min(head_degree, tail_degree)
        low = degree < self.degree_threshold
        edges = seed_edges[low]
        if self.exclude == "reverse_id" and self.reverse_eids is not None:
            return torch.cat([edges, self.reverse_eids[edges.long()]])
        return edges
