"""CappedNeighborSampler (parity: /root/reference/python/dgl/dataloading/capped_neighbor_sampler.py):
neighbour sampling that stops expanding node types once a per-type cap is reached."""
from __future__ import annotations

from collections import defaultdict

import torch

from ..base import NID
from ..sampling import sample_neighbors
from ..transforms import to_block
from .base import Sampler, set_edge_lazy_features, set_node_lazy_features


class CappedNeighborSampler(Sampler):
    def __init__(self, fanouts, fixed_k=5, upsample_rare_types=False, prefetch_node_feats=None, prefetch_edge_feats=None):
        super().__init__()
        self.fanouts = fanouts
        self.fixed_k = fixed_k
        self.upsample_rare_types = upsample_rare_types
        self.prefetch_node_feats = prefetch_node_feats
        self.prefetch_edge_feats = prefetch_edge_feats

    def sample(self, g, indices):
        seed = indices
        if not isinstance(seed, dict):
            seed = {g.ntypes[0]: seed}
        last = {k: v for k, v in seed.items()}
        allnodes = defaultdict(list)
        for k, v in seed.items():
            allnodes[k].append(v)
        for fanout in self.fanouts:
            fr = sample_neighbors(g, last, fanout, copy_ndata=False, copy_edata=False)
            nxt = defaultdict(list)
            for c in fr.canonical_etypes:
                u, _ = fr.all_edges(etype=c)
                if u.numel():
                    cap = self.fixed_k * sum(x.numel() for x in last.values())
                    nxt[c[0]].append(u.unique()[:cap])
            last = {k: torch.cat(v).unique() for k, v in nxt.items()}
            for k, v in last.items():
                allnodes[k].append(v)
        nodes = {k: torch.cat(v).unique() for k, v in allnodes.items()}
        sg = g.subgraph(nodes if len(g.ntypes) > 1 else nodes[g.ntypes[0]], relabel_nodes=True)
        set_node_lazy_features(sg, self.prefetch_node_feats)
        set_edge_lazy_features(sg, self.prefetch_edge_feats)
        return sg
