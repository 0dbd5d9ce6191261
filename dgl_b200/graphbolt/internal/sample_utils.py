"""Compaction helpers used by the samplers (parity: graphbolt/internal/sample_utils.py). The id-dedup kernels live in
``graphbolt/utils.py`` (native order-preserving unique on both devices); this module adds the temporal variant and
re-exports the rest under the reference's path."""
from __future__ import annotations

import torch

from ..utils import compact_csc_format, unique_and_compact, unique_and_compact_csc_formats  # noqa: F401

__all__ = ["unique_and_compact", "compact_temporal_nodes", "unique_and_compact_csc_formats", "compact_csc_format"]


def compact_temporal_nodes(nodes, nodes_timestamp):
    """Concatenate node lists *without* deduplication (a node seen at two timestamps is two temporal nodes) and number
    them consecutively. ``nodes`` / ``nodes_timestamp``: list of tensors, or {type: list of tensors}.
    Returns (all nodes, all timestamps, per-input compacted id tensors)."""
    def one(parts, stamps):
        sizes = [int(p.shape[0]) for p in parts]
        allp, allt = torch.cat(parts), torch.cat(stamps)
        ids = torch.arange(allp.numel(), dtype=allp.dtype, device=allp.device)
        return allp, allt, list(ids.split(sizes))

    if isinstance(nodes, dict):
        out_n, out_t, out_c = {}, {}, {}
        for k in nodes:
            out_n[k], out_t[k], out_c[k] = one(nodes[k], nodes_timestamp[k])
        return out_n, out_t, out_c
    return one(nodes, nodes_timestamp)
