"""ShaDow-GNN k-hop sampler (parity: /root/reference/python/dgl/dataloading/shadow.py)."""
from __future__ import annotations

from ..base import NID
from ..sampling import sample_neighbors
from ..transforms import compact_graphs
from .base import Sampler, set_edge_lazy_features, set_node_lazy_features


class ShaDowKHopSampler(Sampler):
    def __init__(self, fanouts, replace=False, prob=None, prefetch_node_feats=None, prefetch_edge_feats=None,
                 output_device=None):
        super().__init__()
        self.fanouts = fanouts
        self.replace = replace
        self.prob = prob
        self.prefetch_node_feats = prefetch_node_feats
        self.prefetch_edge_feats = prefetch_edge_feats
        self.output_device = output_device

    def sample(self, g, seed_nodes, exclude_eids=None):
        output_nodes = seed_nodes
        for fanout in reversed(self.fanouts):
            frontier = sample_neighbors(g, seed_nodes, fanout, output_device=self.output_device, replace=self.replace,
                                        prob=self.prob, exclude_edges=exclude_eids)
            block = compact_graphs(frontier, seed_nodes)
            seed_nodes = block.ndata[NID]
        subg = g.subgraph(seed_nodes, relabel_nodes=True, output_device=self.output_device)
        if exclude_eids is not None:
            from ..base import EID
            import torch
            ex = exclude_eids if not isinstance(exclude_eids, dict) else None
            if ex is not None:
                keep = ~torch.isin(subg.edata[EID], ex.to(subg.device))
                subg = subg.edge_subgraph(keep.nonzero().squeeze(1), relabel_nodes=False)
        set_node_lazy_features(subg, self.prefetch_node_feats)
        set_edge_lazy_features(subg, self.prefetch_edge_feats)
        return seed_nodes, output_nodes, subg
