"""Multi-node / multi-process distributed graph store and training helpers
(parity: /root/reference/python/dgl/distributed/__init__.py)."""
from .dist_graph import (DistDataLoader, DistEdgeDataLoader, DistEmbedding, DistGraph, DistNodeDataLoader,  # noqa: F401
                         DistSparseAdagrad, DistSparseGradOptimizer, DistTensor, ServerState, edge_split, exit_client,
                         in_subgraph, initialize, is_initialized, node_split, sample_neighbors, DistGraphServer,
                         DistSparseAdam, find_edges, get_num_client, get_rank, sample_etype_neighbors)
from .kvstore import KVClient, KVServer, PartitionPolicy  # noqa: F401
from . import kvstore, optim, role  # noqa: F401
from .partition import (GraphPartitionBook, RangePartitionBook, load_partition, load_partition_book,  # noqa: F401
                        load_partition_feats, partition_graph)
from . import rpc  # noqa: F401
from .rpc import DistConnectError  # noqa: F401
