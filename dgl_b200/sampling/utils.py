"""Sampling utilities (parity: /root/reference/python/dgl/sampling/utils.py)."""
from __future__ import annotations

from collections.abc import Mapping

import torch

from ..base import EID
from ..utils.filter import Filter

__all__ = ["EidExcluder"]


class EidExcluder:
    """Post-filter for sampled frontiers: drops the edges whose *parent* edge id (``edata[dgl.EID]``) is in
    ``exclude_eids`` (tensor, or ``{etype: tensor}`` for heterographs) and re-points ``dgl.EID`` at the parent ids.
    ``weights`` (a one-element list holding the per-edge tensor, as LABOR passes it) is filtered alongside.
    The membership test is the sorted-set :class:`~dgl_b200.utils.Filter` on either device."""

    def __init__(self, exclude_eids):
        if isinstance(exclude_eids, Mapping):
            self._filter = {k: Filter(v) for k, v in exclude_eids.items()}
        else:
            self._filter = Filter(exclude_eids)

    def _drop(self, frontier, positions, etype, weights):
        from ..transforms import remove_edges
        parent = frontier.edges[etype].data[EID] if etype is not None else frontier.edata[EID]
        if positions.numel() == 0:
            return frontier
        keep_weights = weights is not None and etype is None and weights[0].shape[0] == frontier.num_edges()
        out = remove_edges(frontier, positions, etype=etype, store_ids=True)
        kept = out.edges[etype].data[EID] if etype is not None else out.edata[EID]
        if keep_weights:
            weights[0] = weights[0][kept.long()]
        if etype is not None:
            out.edges[etype].data[EID] = parent[kept.long()]
        else:
            out.edata[EID] = parent[kept.long()]
        return out

    def __call__(self, frontier, weights=None):
        if not isinstance(self._filter, Mapping):
            hit = self._filter.find_included_indices(frontier.edata[EID])
            return self._drop(frontier, hit.to(frontier.idtype), None, weights)
        for k, flt in self._filter.items():
            c = frontier.to_canonical_etype(k)
            hit = flt.find_included_indices(frontier.edges[c].data[EID])
            frontier = self._drop(frontier, hit.to(frontier.idtype), c, None)
        return frontier
