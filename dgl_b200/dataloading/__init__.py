"""Mini-batch data loading (``dgl.dataloading``).

Parity: /root/reference/python/dgl/dataloading/ (base.py Sampler/BlockSampler/as_edge_prediction_sampler,
neighbor_sampler.py, labor_sampler.py, cluster_gcn.py, graphsaint.py, shadow.py, negative_sampler.py,
dataloader.py DataLoader :788 / GraphDataLoader :1376 / DDPTensorizedDataset :255).
"""
from .base import (BlockSampler, EdgePredictionSampler, Sampler, as_edge_prediction_sampler, assign_lazy_features,  # noqa: F401
                   set_edge_lazy_features, set_node_lazy_features, set_src_lazy_features, set_dst_lazy_features)
from .cluster_gcn import ClusterGCNSampler  # noqa: F401
from .dataloader import DataLoader, DistNodeDataLoaderStub, GraphDataLoader  # noqa: F401
from .graphsaint import SAINTSampler  # noqa: F401
from .labor_sampler import LaborSampler  # noqa: F401
from . import negative_sampler  # noqa: F401
from .negative_sampler import GlobalUniform, PerSourceUniform, Uniform  # noqa: F401
from .neighbor_sampler import MultiLayerFullNeighborSampler, MultiLayerNeighborSampler, NeighborSampler  # noqa: F401
from .shadow import ShaDowKHopSampler  # noqa: F401
from .capped_neighbor_sampler import CappedNeighborSampler  # noqa: F401
from .spot_target import SpotTarget  # noqa: F401

from .dataloader import GraphCollator  # noqa: F401,E402
from .tensorized import (CollateWrapper, DDPTensorizedDataset, EdgeCollator, NodeCollator, TensorizedDataset,  # noqa: F401,E402
                         WorkerInitWrapper, create_tensorized_dataset, find_exclude_eids,
                         remove_parent_storage_columns, restore_parent_storage_columns)


def __getattr__(name):
    # distributed loaders are also reachable from here, as in the reference's dataloading/dist_dataloader.py
    if name in ("DistDataLoader", "DistNodeDataLoader", "DistEdgeDataLoader"):
        from ..distributed import dist_graph
        return getattr(dist_graph, name)
    raise AttributeError(name)

from .graphed import CapacityBlock, GraphedTrainer  # noqa: F401,E402

from ..frame import LazyFeature  # noqa: F401,E402
from ..sampling.utils import EidExcluder  # noqa: F401,E402
