"""Abstract sampling graph (parity: /root/reference/python/dgl/graphbolt/sampling_graph.py); the implementation is
impl/fused_csc_sampling_graph.py."""
from __future__ import annotations

__all__ = ["SamplingGraph"]


class SamplingGraph:
    """Abstract sampling graph (reference sampling_graph.py)."""

    @property
    def num_nodes(self):
        raise NotImplementedError

    @property
    def num_edges(self):
        raise NotImplementedError

    def copy_to_shared_memory(self, shared_memory_name):
        raise NotImplementedError

    def to(self, device):
        raise NotImplementedError
