"""InSubgraphSampler (parity: /root/reference/python/dgl/graphbolt/impl/in_subgraph_sampler.py)."""
from __future__ import annotations

from ..datapipes import SubgraphSampler
from ..minibatch import SampledSubgraphImpl
from ..utils import unique_and_compact_csc_formats


class InSubgraphSampler(SubgraphSampler):
    def __init__(self, datapipe, graph):
        super().__init__(datapipe)
        self.graph = graph

    def sample_subgraphs(self, seeds, seeds_timestamp=None):
        sg = self.graph.in_subgraph(seeds)
        uniq, compacted = unique_and_compact_csc_formats(sg.sampled_csc, seeds)
        out = SampledSubgraphImpl(sampled_csc=compacted, original_column_node_ids=seeds, original_row_node_ids=uniq,
                                  original_edge_ids=sg.original_edge_ids)
        return uniq, [out]
