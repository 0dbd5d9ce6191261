"""Uniform negative sampler datapipe (parity: /root/reference/python/dgl/graphbolt/impl/uniform_negative_sampler.py,
negative_sampler.py)."""
from __future__ import annotations

import torch

from ..base import etype_str_to_tuple
from ..datapipes import MiniBatchTransformer


class NegativeSampler(MiniBatchTransformer):
    def __init__(self, datapipe, negative_ratio):
        super().__init__(datapipe, self._sample)
        assert negative_ratio > 0, "Negative_ratio should be positive Integer."
        self.negative_ratio = negative_ratio

    def _sample(self, minibatch):
        seeds = minibatch.seeds
        if isinstance(seeds, dict):
            if minibatch.indexes is None:
                minibatch.indexes = {}
            if minibatch.labels is None:
                minibatch.labels = {}
            for etype, pos_pairs in seeds.items():
                s, lab, idx = self._collate(pos_pairs, etype)
                minibatch.seeds[etype], minibatch.labels[etype], minibatch.indexes[etype] = s, lab, idx
        else:
            minibatch.seeds, minibatch.labels, minibatch.indexes = self._collate(seeds, None)
        return minibatch

    def _collate(self, pos_pairs, etype):
        neg = self._sample_with_etype(pos_pairs, etype)
        seeds = torch.cat([pos_pairs, neg])
        pos_num = pos_pairs.shape[0]
        labels = torch.cat([torch.ones(pos_num, device=seeds.device), torch.zeros(neg.shape[0], device=seeds.device)])
        pos_idx = torch.arange(pos_num, device=seeds.device)
        indexes = torch.cat([pos_idx, pos_idx.repeat_interleave(self.negative_ratio)])
        return seeds, labels, indexes

    def _sample_with_etype(self, node_pairs, etype=None):
        raise NotImplementedError


class UniformNegativeSampler(NegativeSampler):
    """For every positive (u, v) draw ``negative_ratio`` pairs (u, v') with v' uniform over destination nodes."""

    def __init__(self, datapipe, graph, negative_ratio):
        super().__init__(datapipe, negative_ratio)
        self.graph = graph

    def _sample_with_etype(self, node_pairs, etype=None):
        if etype is not None:
            _, _, d = etype_str_to_tuple(etype)
            max_id = self.graph.num_nodes[d]
        else:
            max_id = self.graph.total_num_nodes
        src = node_pairs[:, 0].repeat_interleave(self.negative_ratio)
        dst = torch.randint(0, max_id, (src.shape[0],), dtype=node_pairs.dtype, device=node_pairs.device)
        return torch.stack([src, dst], 1)
