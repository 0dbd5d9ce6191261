"""Reference module path ``dgl.distributed.graph_partition_book``: the names below are implemented in
dgl_b200.distributed, dgl_b200.distributed.compat, dgl_b200.distributed.rpc_api ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.distributed', 'dgl_b200.distributed.compat', 'dgl_b200.distributed.rpc_api', 'dgl_b200.distributed.dist_graph', 'dgl_b200.distributed.kvstore', 'dgl_b200.distributed.partition', 'dgl_b200.distributed.optim', 'dgl_b200.optim', 'dgl_b200.nn', 'dgl_b200'],
                                          ['EdgePartitionPolicy', 'GraphPartitionBook', 'HeteroDataName', 'NodePartitionPolicy', 'PartitionPolicy', 'RangePartitionBook', 'get_node_partition_from_book', 'get_shared_mem_partition_book', 'parse_hetero_data_name'])
