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


# Names the reference's ``graphbolt.impl`` namespace also carries (its submodules star-import one another); they live in
# the parent package here and are resolved on first use, which keeps ``impl`` importable before the parent finishes.
_PARENT_NAMES = {
    "CPUFeatureCache": ("dgl_b200.graphbolt.impl.caches", "FeatureCache"),
    "GPUFeatureCache": ("dgl_b200.graphbolt.impl.caches", "GPUCache"),
    "SampledSubgraphImpl": ("dgl_b200.graphbolt.minibatch", None), "ItemSet": ("dgl_b200.graphbolt.itemset", None),
    "HeteroItemSet": ("dgl_b200.graphbolt.itemset", None), "Dataset": ("dgl_b200.graphbolt.dataset", None),
    "Task": ("dgl_b200.graphbolt.dataset", None), "SamplingGraph": ("dgl_b200.graphbolt.sampling_graph", None),
    "LegacyDataset": ("dgl_b200.graphbolt.impl.legacy_dataset", None),
    "etype_tuple_to_str": ("dgl_b200.graphbolt.base", None),
    "load_from_shared_memory": ("dgl_b200.graphbolt.impl.fused_csc_sampling_graph", None),
    "preprocess_ondisk_dataset": ("dgl_b200.graphbolt.compat", None), "SamplePerLayer": ("dgl_b200.graphbolt.compat", None),
    "FetchInsubgraphData": ("dgl_b200.graphbolt.compat", None),
    "CombineCachedAndFetchedInSubgraph": ("dgl_b200.graphbolt.compat", None),
}
for _n in ("OnDiskFeatureData", "OnDiskFeatureDataDomain", "OnDiskFeatureDataFormat", "OnDiskGraphTopology",
           "OnDiskGraphTopologyType", "OnDiskMetaData", "OnDiskTVTSet", "OnDiskTVTSetData", "OnDiskTaskData"):
    _PARENT_NAMES[_n] = ("dgl_b200.graphbolt.impl.ondisk_metadata", None)


def __getattr__(name):
    hit = _PARENT_NAMES.get(name)
    if hit is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    import importlib
    return getattr(importlib.import_module(hit[0]), hit[1] or name)
