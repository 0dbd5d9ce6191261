"""``SubgraphSampler``: turns the seeds of a mini-batch into unique seed nodes (+ compacted seeds for link / hyperlink
tasks) and delegates to ``sample_subgraphs``.

Parity: /root/reference/python/dgl/graphbolt/subgraph_sampler.py:27-310. The cooperative (multi-GPU) variant of the
reference's four exchange stages lives in impl/cooperative.py (``ExchangePlan``) and is driven from
impl/neighbor_sampler.py."""
from __future__ import annotations

import torch

from .base import etype_str_to_tuple
from .minibatch_transformer import MiniBatchTransformer

__all__ = ["SubgraphSampler"]


class SubgraphSampler(MiniBatchTransformer):
    """Base: seeds -> unique seed nodes (+ compacted seeds for edge/hyperlink tasks) -> sample_subgraphs()."""

    def __init__(self, datapipe, *args, **kwargs):
        super().__init__(datapipe, self._sample)
        self._args, self._kwargs = args, kwargs

    @staticmethod
    def _seeds_preprocess(minibatch):
        from .utils import unique_and_compact
        seeds = minibatch.seeds
        if isinstance(seeds, dict):
            nodes = {}
            compact_in = {}
            for et, s in seeds.items():
                if s.dim() == 1:
                    nodes.setdefault(et, []).append(s)
                else:
                    src_t, _, dst_t = etype_str_to_tuple(et)
                    nodes.setdefault(src_t, []).append(s[:, 0])
                    nodes.setdefault(dst_t, []).append(s[:, 1])
            uniq, comp = unique_and_compact(nodes)
            compacted = {}
            cursor = {k: 0 for k in nodes}
            for et, s in seeds.items():
                if s.dim() == 1:
                    cursor[et] += 1
                    continue
                src_t, _, dst_t = etype_str_to_tuple(et)
                a = comp[src_t][cursor[src_t]]; cursor[src_t] += 1
                b = comp[dst_t][cursor[dst_t]]; cursor[dst_t] += 1
                compacted[et] = torch.stack([a, b], 1)
            return uniq, (compacted or None)
        if seeds.dim() == 1:
            uniq, _ = unique_and_compact([seeds])
            return seeds, None  # node seeds are used as-is (order preserved)
        cols = [seeds[:, i] for i in range(seeds.shape[1])]
        uniq, comp = unique_and_compact(cols)
        return uniq, torch.stack(comp, 1)

    def _sample(self, minibatch):
        seeds, compacted = self._seeds_preprocess(minibatch)
        minibatch.input_nodes, minibatch.sampled_subgraphs = self.sample_subgraphs(seeds)
        minibatch.compacted_seeds = compacted
        return minibatch

    def sample_subgraphs(self, seeds, seeds_timestamp=None):
        raise NotImplementedError


def __getattr__(name):
    # remaining names of the reference module of this path (markers, helpers) are defined in graphbolt/compat.py
    if name.startswith("__"):
        raise AttributeError(name)
    from . import compat
    try:
        return getattr(compat, name)
    except AttributeError:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}") from None
