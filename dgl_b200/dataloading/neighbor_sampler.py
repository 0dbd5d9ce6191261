"""NeighborSampler (API parity: /root/reference/python/dgl/dataloading/neighbor_sampler.py:152-206). On a GPU-resident
homogeneous graph the layers are sampled by the one-synchronisation pipeline of ``sampling/fused_blocks.py``; everything
else (heterographs, weights, replacement, CPU) takes the layer-by-layer ``sample_neighbors`` + ``to_block`` route."""
from __future__ import annotations

import torch

from ..base import EID, NID
from ..sampling import sample_neighbors
from ..transforms import to_block
from .base import BlockSampler


class NeighborSampler(BlockSampler):
    """Multi-layer neighbour sampler producing MFGs; ``fanouts[i]`` is the fan-out of layer i
    (int or dict per edge type, -1 = all)."""

    def __init__(self, fanouts, edge_dir="in", prob=None, mask=None, replace=False, prefetch_node_feats=None,
                 prefetch_labels=None, prefetch_edge_feats=None, output_device=None, fused=True):
        super().__init__(prefetch_node_feats=prefetch_node_feats, prefetch_labels=prefetch_labels,
                         prefetch_edge_feats=prefetch_edge_feats, output_device=output_device)
        self.fanouts = fanouts
        self.edge_dir = edge_dir
        if mask is not None and prob is not None:
            raise ValueError("Mask and probability arguments are mutually exclusive. "
                             "Consider multiplying the probability with the mask to achieve the same goal.")
        self.prob = prob or mask
        self.replace = replace
        self.fused = fused
        self.mapping = {}
        self._fast = None

    def sample_blocks(self, g, seed_nodes, exclude_eids=None):
        # GPU fast path: all layers enqueued back to back on capacity buffers, ONE host synchronisation per batch
        # (sampling/fused_blocks.py); blocks carry their ELL form for the tcgen05 SAGE layer
        if self.fused and torch.is_tensor(seed_nodes) and not isinstance(self.fanouts, dict):
            from ..sampling import fused_blocks
            if fused_blocks.usable(g, self.fanouts, self.prob, self.replace, self.edge_dir, exclude_eids) and \
                    seed_nodes.device == g.device and self.output_device in (None, g.device):
                if self._fast is None:
                    self._fast = fused_blocks.FusedBlockSampler(self.fanouts)
                return self._fast(g, seed_nodes)
        output_nodes = seed_nodes
        blocks = []
        for fanout in reversed(self.fanouts):
            frontier = sample_neighbors(g, seed_nodes, fanout, edge_dir=self.edge_dir, prob=self.prob,
                                        replace=self.replace, output_device=self.output_device,
                                        exclude_edges=exclude_eids, copy_ndata=False, copy_edata=False)
            eid = frontier.edata[EID]
            block = to_block(frontier, seed_nodes)
            block.edata[EID] = eid
            seed_nodes = block.srcdata[NID]
            blocks.insert(0, block)
        return seed_nodes, output_nodes, blocks


MultiLayerNeighborSampler = NeighborSampler


class MultiLayerFullNeighborSampler(NeighborSampler):
    def __init__(self, num_layers, **kwargs):
        super().__init__([-1] * num_layers, **kwargs)
