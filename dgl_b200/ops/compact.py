"""Id compaction / relabel helpers (CPU native table or sm_100a hash kernel).

Parity: /root/reference/src/runtime/cuda/cuda_hashtable.cu (ordered table),
src/graph/transform/cuda/cuda_map_edges.cuh, graphbolt unique_and_compact.
"""
from __future__ import annotations

import torch

from .._native import loader
from ..base import id_code


def map_ids(universe, ids):
    """local[i] = position of ids[i] in ``universe`` (-1 if absent)."""
    idt = ids.dtype
    out = torch.empty_like(ids)
    if ids.numel() == 0:
        return out
    universe = universe.to(idt).contiguous()
    ids = ids.contiguous()
    if ids.is_cuda:
        from .cuda_struct import cuda_map_ids
        return cuda_map_ids(universe, ids)
    loader.cpu_lib().fn("dglb200_cpu_map_ids", "iplplp")(id_code(idt), universe.data_ptr(), universe.shape[0],
                                                         ids.data_ptr(), ids.shape[0], out.data_ptr())
    return out


def cuda_unique_compact(first, ids):
    from .cuda_struct import cuda_unique_compact as f
    return f(first, ids)
