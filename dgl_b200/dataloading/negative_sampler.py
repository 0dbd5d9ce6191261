"""Negative samplers (parity: /root/reference/python/dgl/dataloading/negative_sampler.py)."""
from __future__ import annotations

from collections.abc import Mapping

import torch


class _BaseNegativeSampler:
    def _generate(self, g, eids, canonical_etype):
        raise NotImplementedError

    def __call__(self, g, eids):
        if isinstance(eids, Mapping):
            eids = {g.to_canonical_etype(k): v for k, v in eids.items()}
            return {k: self._generate(g, v, k) for k, v in eids.items()}
        return self._generate(g, eids, g.canonical_etypes[0])


class PerSourceUniform(_BaseNegativeSampler):
    """k uniformly drawn destination nodes per source of every positive edge."""

    def __init__(self, k):
        self.k = k

    def _generate(self, g, eids, canonical_etype):
        _, _, vtype = canonical_etype
        src, _ = g.find_edges(eids, etype=canonical_etype)
        src = src.repeat_interleave(self.k)
        dst = torch.randint(0, g.num_nodes(vtype), (src.shape[0],), dtype=src.dtype, device=src.device)
        return src, dst


Uniform = PerSourceUniform


class GlobalUniform(_BaseNegativeSampler):
    """k non-edges per positive edge drawn from the whole graph."""

    def __init__(self, k, exclude_self_loops=True, replace=False):
        self.k = k
        self.exclude_self_loops = exclude_self_loops
        self.replace = replace

    def _generate(self, g, eids, canonical_etype):
        from ..sampling import global_uniform_negative_sampling
        return global_uniform_negative_sampling(g, len(eids) * self.k, self.exclude_self_loops, self.replace,
                                                canonical_etype)
