"""GraphSAINT samplers (parity: /root/reference/python/dgl/dataloading/graphsaint.py)."""
from __future__ import annotations

import torch

from ..base import DGLError
from ..sampling import random_walk
from .base import Sampler, set_edge_lazy_features, set_node_lazy_features


class SAINTSampler(Sampler):
    """mode in {'node', 'edge', 'walk'}; budget = #nodes | #edges | (num_roots, walk_length)."""

    def __init__(self, mode, budget, cache=True, prefetch_ndata=None, prefetch_edata=None, output_device="cpu"):
        super().__init__()
        self.budget = budget
        if mode == "node":
            self.sampler = self.node_sampler
        elif mode == "edge":
            self.sampler = self.edge_sampler
        elif mode == "walk":
            self.sampler = self.walk_sampler
        else:
            raise DGLError(f"Expect mode to be 'node', 'edge' or 'walk', got {mode}.")
        self.cache = cache
        self.prob = None
        self.prefetch_ndata = prefetch_ndata or []
        self.prefetch_edata = prefetch_edata or []
        self.output_device = output_device

    def node_sampler(self, g):
        if self.cache and self.prob is not None:
            prob = self.prob
        else:
            prob = g.out_degrees().float().clamp(min=1)
            if self.cache:
                self.prob = prob
        return torch.multinomial(prob, num_samples=self.budget, replacement=True).unique().type(g.idtype)

    def edge_sampler(self, g):
        src, dst = g.edges()
        if self.cache and self.prob is not None:
            prob = self.prob
        else:
            in_deg = g.in_degrees().float().clamp(min=1)
            out_deg = g.out_degrees().float().clamp(min=1)
            prob = 1.0 / in_deg[dst.long()] + 1.0 / out_deg[src.long()]
            prob /= prob.sum()
            if self.cache:
                self.prob = prob
        sampled_edges = torch.unique(torch.multinomial(prob, num_samples=self.budget, replacement=True))
        return torch.cat([src[sampled_edges], dst[sampled_edges]]).unique().type(g.idtype)

    def walk_sampler(self, g):
        num_roots, walk_length = self.budget
        sampled_roots = torch.randint(0, g.num_nodes(), (num_roots,), device=g.device)
        traces, types = random_walk(g, nodes=sampled_roots, length=walk_length)
        sampled_nodes = traces[traces >= 0].unique()
        return sampled_nodes.type(g.idtype)

    def sample(self, g, indices):
        node_ids = self.sampler(g)
        sg = g.subgraph(node_ids, relabel_nodes=True, output_device=self.output_device)
        set_node_lazy_features(sg, self.prefetch_ndata)
        set_edge_lazy_features(sg, self.prefetch_edata)
        return sg
