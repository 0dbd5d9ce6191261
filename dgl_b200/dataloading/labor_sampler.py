"""LaborSampler (parity: /root/reference/python/dgl/dataloading/labor_sampler.py)."""
from __future__ import annotations

import torch

from .. import random as dgl_random
from ..base import EID, NID
from ..sampling import sample_labors
from ..transforms import to_block
from .base import BlockSampler


class LaborSampler(BlockSampler):
    def __init__(self, fanouts, edge_dir="in", prob=None, importance_sampling=0, layer_dependency=False,
                 batch_dependency=1, prefetch_node_feats=None, prefetch_labels=None, prefetch_edge_feats=None,
                 output_device=None):
        super().__init__(prefetch_node_feats=prefetch_node_feats, prefetch_labels=prefetch_labels,
                         prefetch_edge_feats=prefetch_edge_feats, output_device=output_device)
        self.fanouts = fanouts
        self.edge_dir = edge_dir
        self.prob = prob
        self.importance_sampling = importance_sampling
        self.layer_dependency = layer_dependency
        self.cnt = 0
        self.batch_dependency = batch_dependency
        self.random_seed = dgl_random.next_seed()

    def set_seed(self, random_seed=None):
        """Re-seed the shared per-vertex variates (reference labor_sampler.py:set_seed). With ``batch_dependency`` > 1
        consecutive batches interpolate between successive seeds; calling this restarts that schedule so that every
        rank / worker given the same ``random_seed`` draws identical neighbourhoods."""
        self.random_seed = dgl_random.next_seed() if random_seed is None else int(random_seed)
        self.cnt = 0

    def sample_blocks(self, g, seed_nodes, exclude_eids=None):
        output_nodes = seed_nodes
        blocks = []
        self.cnt += 1
        t = (self.cnt % self.batch_dependency) / self.batch_dependency if self.batch_dependency > 1 else 0.0
        base = self.random_seed + (self.cnt // max(self.batch_dependency, 1)) * 7919
        for i, fanout in enumerate(reversed(self.fanouts)):
            seed_i = base if self.layer_dependency else base + i + 1
            frontier, importances = sample_labors(g, seed_nodes, fanout, self.edge_dir, self.prob,
                                                  importance_sampling=self.importance_sampling, random_seed=seed_i,
                                                  seed2_contribution=t, output_device=self.output_device,
                                                  exclude_edges=exclude_eids)
            eid = frontier.edata[EID]
            block = to_block(frontier, seed_nodes)
            block.edata[EID] = eid
            if len(g.canonical_etypes) > 1:
                for etype, importance in zip(g.canonical_etypes, importances):
                    if importance.shape[0] == block.num_edges(etype):
                        block.edata["edge_weights"] = {etype: importance}
            elif importances[0].shape[0] == block.num_edges():
                block.edata["edge_weights"] = importances[0]
            seed_nodes = block.srcdata[NID]
            blocks.insert(0, block)
        return seed_nodes, output_nodes, blocks
