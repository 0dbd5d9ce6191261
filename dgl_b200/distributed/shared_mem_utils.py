"""Reference module path ``dgl.distributed.shared_mem_utils``: the names below are implemented in
dgl_b200.distributed, dgl_b200.distributed.compat, dgl_b200.distributed.rpc_api ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.distributed', 'dgl_b200.distributed.compat', 'dgl_b200.distributed.rpc_api', 'dgl_b200.distributed.dist_graph', 'dgl_b200.distributed.kvstore', 'dgl_b200.distributed.partition', 'dgl_b200.distributed.optim', 'dgl_b200.optim', 'dgl_b200.nn', 'dgl_b200'],
                                          [])
