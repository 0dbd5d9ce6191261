"""Negative-edge samplers for edge-prediction loaders.

A negative sampler is a callable ``(graph, seed edge ids) -> (src, dst)`` — or a dict of such pairs keyed by canonical
edge type when the seeds are a dict. :class:`PerSourceUniform` (alias ``Uniform``) corrupts the destination of every
positive edge ``k`` times; :class:`GlobalUniform` draws node pairs from the whole graph that are verified non-edges
(native rejection sampler, ``sampling/negative.py``). Role of the reference's dataloading/negative_sampler.py."""
from __future__ import annotations

from collections.abc import Mapping

import torch


class _BaseNegativeSampler:
    """Dispatch over edge types; subclasses implement :meth:`pairs_for` for one relation."""

    def pairs_for(self, g, eids, canonical_etype):
        raise NotImplementedError

    # name used by subclasses written against the reference's base class
    def _generate(self, g, eids, canonical_etype):
        return self.pairs_for(g, eids, canonical_etype)

    def __call__(self, g, eids):
        if not isinstance(eids, Mapping):
            return self._generate(g, eids, g.canonical_etypes[0])
        out = {}
        for etype, ids in eids.items():
            rel = g.to_canonical_etype(etype)
            out[rel] = self._generate(g, ids, rel)
        return out


class PerSourceUniform(_BaseNegativeSampler):
    """For every positive edge (u, v): ``k`` pairs (u, v') with v' uniform over the destination node type."""

    def __init__(self, k):
        self.k = int(k)

    def pairs_for(self, g, eids, canonical_etype):
        heads = g.find_edges(eids, etype=canonical_etype)[0].repeat_interleave(self.k)
        n_dst = g.num_nodes(canonical_etype[2])
        tails = torch.randint(0, n_dst, heads.shape, dtype=heads.dtype, device=heads.device)
        return heads, tails


Uniform = PerSourceUniform


class GlobalUniform(_BaseNegativeSampler):
    """``k`` verified non-edges per positive edge, drawn uniformly over all node pairs of the relation; fewer may be
    returned on dense graphs. ``exclude_self_loops`` / ``replace`` are forwarded to
    :func:`dgl_b200.sampling.global_uniform_negative_sampling`."""

    def __init__(self, k, exclude_self_loops=True, replace=False):
        self.k, self.exclude_self_loops, self.replace = int(k), exclude_self_loops, replace

    def pairs_for(self, g, eids, canonical_etype):
        from ..sampling import global_uniform_negative_sampling as draw
        return draw(g, len(eids) * self.k, self.exclude_self_loops, self.replace, canonical_etype)
