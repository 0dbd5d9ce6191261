"""GraphBolt NeighborSampler / LayerNeighborSampler datapipes (parity:
/root/reference/python/dgl/graphbolt/impl/neighbor_sampler.py:745-1098)."""
from __future__ import annotations

import torch

from ..datapipes import SubgraphSampler
from ..minibatch import SampledSubgraphImpl
from ..utils import unique_and_compact_csc_formats


class NeighborSampler(SubgraphSampler):
    """Multi-layer neighbour sampling on a FusedCSCSamplingGraph; ``fanouts[i]`` = fan-out(s) of layer i
    (an int/tensor, or one entry per edge type)."""

    def __init__(self, datapipe, graph, fanouts, replace=False, prob_name=None, deduplicate=True, overlap_fetch=False,
                 num_gpu_cached_edges=0, gpu_cache_threshold=1, cooperative=False, asynchronous=False, layer=False,
                 batch_dependency=1, layer_dependency=False):
        super().__init__(datapipe)
        self.graph = graph
        self.fanouts = [torch.as_tensor(f).reshape(-1) for f in fanouts]
        self.replace = replace
        self.prob_name = prob_name
        self.deduplicate = deduplicate
        self.layer = layer
        self.batch_dependency = batch_dependency
        self.layer_dependency = layer_dependency
        self.cooperative = bool(cooperative)
        self._cnt = 0
        self._fast = None

    def _sample_one(self, seeds, fanout, layer_idx):
        if self.layer:
            from ... import random as dgl_random
            self_seed = getattr(self, "_base_seed", None)
            if self_seed is None:
                self._base_seed = self_seed = dgl_random.next_seed()
            t = (self._cnt % self.batch_dependency) / self.batch_dependency if self.batch_dependency > 1 else 0.0
            rs = self_seed + (self._cnt // max(self.batch_dependency, 1)) * 7919 + (0 if self.layer_dependency else layer_idx + 1)
            return self.graph.sample_layer_neighbors(seeds, fanout, self.replace, self.prob_name, random_seed=rs,
                                                     seed2_contribution=t)
        return self.graph.sample_neighbors(seeds, fanout, self.replace, self.prob_name)

    def _fast_usable(self, seeds):
        from ...sampling.fused_blocks import ENABLED
        if not ENABLED[0]:
            return False
        g = self.graph
        if self.layer or self.replace or self.prob_name is not None or not self.deduplicate or not torch.is_tensor(seeds):
            return False
        if g.type_per_edge is not None or g.node_type_offset is not None or not seeds.is_cuda:
            return False
        if not (g.csc_indptr.is_cuda or g.is_pinned()) or g.indices.shape[0] >= 2 ** 31:
            return False
        return all(f.numel() == 1 and 1 <= int(f) <= 64 for f in self.fanouts)

    def _sample_fast(self, seeds):
        """All hops through the one-synchronisation pipeline (sampling/fused_blocks.py); the graph may be device
        resident or host-pinned (the sampling kernel reads pinned CSC arrays over UVA)."""
        from ...sampling.fused_blocks import FusedBlockSampler
        g = self.graph
        if self._fast is None:
            self._fast = FusedBlockSampler([int(f) for f in self.fanouts])
            to32 = lambda t: t if t.dtype == torch.int32 else t.to(torch.int32)  # noqa: E731
            ip, ix = to32(g.csc_indptr), to32(g.indices)
            if g.is_pinned() and not ip.is_pinned():
                from ...utils import pin_memory_inplace
                pin_memory_inplace(ip), pin_memory_inplace(ix)
            self._fast_csc = (ip.contiguous(), ix.contiguous())
        ip, ix = self._fast_csc
        idt = seeds.dtype
        subgraphs = []
        from ..base import CSCFormatBase
        for layer in self._fast.sample_layers(ip, ix, seeds, uva=not ip.is_cuda):
            block, ptr, idx, eids = self._fast.layer_to_block(layer, idt, [("_N", "_E", "_N")], ["_N"])
            sg = SampledSubgraphImpl(sampled_csc=CSCFormatBase(ptr, idx), original_column_node_ids=block.dstdata["_ID"],
                                     original_row_node_ids=block.srcdata["_ID"], original_edge_ids=eids)
            sg._block = block                 # MiniBatch.blocks hands this out as is (no re-validation of the structure)
            subgraphs.append(sg)
        return subgraphs[0].original_row_node_ids, subgraphs

    def _sample_cooperative(self, seeds):
        """All ranks sample together: every vertex of every frontier is expanded by the one rank that owns it (hash
        partition of the id space); see impl/cooperative.py. The minibatch's input nodes are this rank's OWNED inputs and
        each subgraph carries the exchange plan ``CooperativeConv`` uses to fetch the rows of its source nodes."""
        from .cooperative import cooperative_sample
        if not torch.is_tensor(seeds) or self.graph.node_type_offset is not None:
            from ...base import DGLError
            raise DGLError("cooperative minibatching is implemented for homogeneous graphs")
        inputs, layers, out_plan = cooperative_sample(
            lambda owned, hop: self._sample_one(owned, self.fanouts[hop], hop).sampled_csc, seeds, len(self.fanouts))
        subgraphs = []
        for lay in layers:
            sg = SampledSubgraphImpl(sampled_csc=lay["csc"], original_column_node_ids=lay["dst"], original_row_node_ids=lay["src"])
            plan = lay["plan"]
            sg._plan = plan
            sg._counts_sent, sg._counts_received, sg._seed_inverse_ids = plan.send_counts, plan.recv_counts, plan.recv_to_owned
            subgraphs.append(sg)
        self._out_plan = out_plan
        return inputs, subgraphs

    def _sample(self, minibatch):
        minibatch = super()._sample(minibatch)
        if self.cooperative:
            minibatch._out_plan = self._out_plan          # routes the model output back to the ranks that asked for it
        return minibatch

    def sample_subgraphs(self, seeds, seeds_timestamp=None):
        subgraphs = []
        self._cnt += 1
        if self.cooperative:
            return self._sample_cooperative(seeds)
        if self._fast_usable(seeds):
            return self._sample_fast(seeds)
        for hop in reversed(range(len(self.fanouts))):
            sg = self._sample_one(seeds, self.fanouts[hop], hop)
            if self.deduplicate:
                uniq, compacted = unique_and_compact_csc_formats(sg.sampled_csc, seeds)
            else:
                from ..utils import compact_csc_format
                uniq, compacted = compact_csc_format(sg.sampled_csc, seeds)
            subgraphs.insert(0, SampledSubgraphImpl(sampled_csc=compacted, original_column_node_ids=seeds,
                                                    original_row_node_ids=uniq, original_edge_ids=sg.original_edge_ids))
            seeds = uniq
        return seeds, subgraphs


class LayerNeighborSampler(NeighborSampler):
    """LABOR layer-neighbour sampler (reference impl/neighbor_sampler.py:LayerNeighborSampler)."""

    def __init__(self, datapipe, graph, fanouts, replace=False, prob_name=None, deduplicate=True, layer_dependency=False,
                 batch_dependency=1, **kw):
        super().__init__(datapipe, graph, fanouts, replace, prob_name, deduplicate, layer=True,
                         batch_dependency=batch_dependency, layer_dependency=layer_dependency)


class TemporalNeighborSampler(SubgraphSampler):
    """Neighbour sampling that only follows edges / neighbours older than the seed's timestamp
    (reference impl/temporal_neighbor_sampler.py). Nodes are NOT deduplicated: the same node reached at two different
    times is two rows; every sampled source inherits the timestamp of the seed it was sampled for."""

    def __init__(self, datapipe, graph, fanouts, replace=False, prob_name=None, node_timestamp_attr_name=None,
                 edge_timestamp_attr_name=None, layer=False):
        super().__init__(datapipe)
        self.graph = graph
        self.fanouts = [torch.as_tensor(f).reshape(-1) for f in fanouts]
        self.replace, self.prob_name = replace, prob_name
        self.node_ts, self.edge_ts = node_timestamp_attr_name, edge_timestamp_attr_name
        self.layer = layer

    def _sample(self, minibatch):
        seeds = minibatch.seeds
        ts = getattr(minibatch, "timestamp", None)
        if ts is None:
            raise ValueError("TemporalNeighborSampler needs a 'timestamp' field in the item set")
        compacted = None
        if seeds.dim() == 2:  # link / hyperlink seeds: every endpoint is looked up at the time of its edge
            k = seeds.shape[1]
            compacted = torch.arange(seeds.numel(), device=seeds.device).reshape(-1, k)
            ts = ts.repeat_interleave(k)
            seeds = seeds.reshape(-1)
        minibatch.input_nodes, minibatch.sampled_subgraphs = self.sample_subgraphs(seeds, ts)
        minibatch.compacted_seeds = compacted
        return minibatch

    def sample_subgraphs(self, seeds, seeds_timestamp=None):
        from ..utils import compact_csc_format
        subgraphs = []
        ts = seeds_timestamp
        for hop in reversed(range(len(self.fanouts))):
            pick = self.graph.temporal_sample_layer_neighbors if self.layer else self.graph.temporal_sample_neighbors
            sg = pick(seeds, ts, self.fanouts[hop], self.replace, None, self.prob_name, self.node_ts, self.edge_ts)
            csc = sg.sampled_csc
            deg = (csc.indptr[1:] - csc.indptr[:-1]).long()
            rows, compacted = compact_csc_format(csc, seeds)
            subgraphs.insert(0, SampledSubgraphImpl(sampled_csc=compacted, original_column_node_ids=seeds,
                                                    original_row_node_ids=rows, original_edge_ids=sg.original_edge_ids))
            ts = torch.cat([ts, torch.repeat_interleave(ts, deg)])
            seeds = rows
        return seeds, subgraphs


class TemporalLayerNeighborSampler(TemporalNeighborSampler):
    """Layer-wise variant: source vertices carry one shared variate per call (impl/temporal.py), so overlapping temporal
    neighbourhoods select overlapping sources."""

    def __init__(self, datapipe, graph, fanouts, replace=False, prob_name=None, node_timestamp_attr_name=None,
                 edge_timestamp_attr_name=None):
        super().__init__(datapipe, graph, fanouts, replace, prob_name, node_timestamp_attr_name, edge_timestamp_attr_name,
                         layer=True)


_COMPAT_NAMES = ("NeighborSamplerImpl", "SamplePerLayer", "CompactPerLayer", "FetchInsubgraphData",
                 "FetchCachedInsubgraphData", "CombineCachedAndFetchedInSubgraph")


def __getattr__(name):
    # per-stage classes of the reference's split pipeline are defined in graphbolt/compat.py
    if name in _COMPAT_NAMES:
        from .. import compat
        return getattr(compat, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
