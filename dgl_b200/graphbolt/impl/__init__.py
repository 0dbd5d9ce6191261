from .fused_csc_sampling_graph import FusedCSCSamplingGraph, from_dglgraph, fused_csc_sampling_graph  # noqa: F401
from .neighbor_sampler import (LayerNeighborSampler, NeighborSampler, TemporalLayerNeighborSampler,  # noqa: F401
                               TemporalNeighborSampler)
from .negative_sampler import UniformNegativeSampler  # noqa: F401
from .in_subgraph_sampler import InSubgraphSampler  # noqa: F401
from .features import (BasicFeatureStore, CPUCachedFeature, DiskBasedFeature, Feature, GPUCachedFeature,  # noqa: F401
                       TorchBasedFeature, TorchBasedFeatureStore, cpu_cached_feature, gpu_cached_feature)
from .caches import CachePolicyKind, FeatureCache, GPUCache, GPUGraphCache, index_select_csc  # noqa: F401
from .datasets import BuiltinDataset, OnDiskDataset, OnDiskTask  # noqa: F401
from .cooperative import CooperativeConv, CooperativeConvFunction  # noqa: F401
