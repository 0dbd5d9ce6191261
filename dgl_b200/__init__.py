"""dgl_b200 — a Blackwell (sm_100a) native graph neural network framework with the
capabilities and API surface of DGL. ``import dgl_b200 as dgl``.

Layer map: native kernels (csrc/cpu OpenMP, csrc/cuda sm_100a) ->
``_sparse_ops`` dispatch -> ``ops`` autograd -> DGLGraph/``core`` message passing ->
``nn``/``optim`` -> ``sampling``/``dataloading``/``graphbolt`` -> ``parallel``/``distributed``.
"""
from . import _pathshim as _pathshim

_pathshim.install()   # reference module paths without a file of their own (see _alias_table.py)

from . import function  # noqa: F401,E402
from .base import ALL, EID, ETYPE, NID, NTYPE, DGLError, dgl_warning  # noqa: F401
from .batch import batch, batch_hetero, slice_batch, unbatch, unbatch_hetero  # noqa: F401
from .convert import (bipartite_from_networkx, bipartite_from_scipy, block_to_graph, create_block,  # noqa: F401
                      from_cugraph, from_networkx, from_scipy, graph, hetero_from_shared_memory, heterograph,
                      rand_bipartite, rand_graph, to_cugraph,
                      to_heterogeneous, to_homogeneous, to_networkx)
from .frame import LazyFeature  # noqa: F401
from .heterograph import DGLBlock, DGLGraph, DGLHeteroGraph  # noqa: F401
from .readout import *  # noqa: F401,F403
from .subgraph import *  # noqa: F401,F403
from .transforms import *  # noqa: F401,F403
from . import ops  # noqa: F401
from . import nn  # noqa: F401
from .version import __version__  # noqa: F401

from .random import seed  # noqa: F401
from . import backend  # noqa: F401,E402
from .logging import enable_verbose_logging  # noqa: F401,E402
from .utils import cpu_quota as _cpu_quota  # noqa: E402

_cpu_quota.apply()  # container-aware thread budget (see utils/cpu_quota.py); DGLB200_CPU_QUOTA=0 opts out

_LAZY = {"sampling", "dataloading", "graphbolt", "sparse", "distributed", "optim", "data", "parallel", "models",
         "utils", "multiprocessing", "cuda", "geometry", "partition", "traversal", "propagate", "storages",
         "distgnn", "mpops", "generators", "merge", "homophily", "label_informativeness", "logging", "global_config",
         "init", "random", "frame", "view", "core", "convert", "batch", "subgraph", "readout", "contrib", "graph_index",
         "heterograph_index", "heterograph", "base", "transforms", "udf", "ndarray", "container"}


def __getattr__(name):
    import importlib
    if name in _LAZY:
        mod = importlib.import_module(f"{__name__}.{name}")
        globals()[name] = mod
        return mod
    # convenience re-exports resolved lazily (heavy modules)
    _late = {
        "save_graphs": ("data.graph_serialize", "save_graphs"), "load_graphs": ("data.graph_serialize", "load_graphs"),
        "metis_partition": ("partition", "metis_partition"),
        "bfs_nodes_generator": ("traversal", "bfs_nodes_generator"),
        "bfs_edges_generator": ("traversal", "bfs_edges_generator"),
        "topological_nodes_generator": ("traversal", "topological_nodes_generator"),
        "dfs_edges_generator": ("traversal", "dfs_edges_generator"),
        "dfs_labeled_edges_generator": ("traversal", "dfs_labeled_edges_generator"),
        "prop_nodes": ("propagate", "prop_nodes"), "prop_edges": ("propagate", "prop_edges"),
        "prop_nodes_bfs": ("propagate", "prop_nodes_bfs"), "prop_nodes_topo": ("propagate", "prop_nodes_topo"),
        "prop_edges_dfs": ("propagate", "prop_edges_dfs"),
        "merge": ("merge", "merge"),
        "node_homophily": ("homophily", "node_homophily"), "edge_homophily": ("homophily", "edge_homophily"),
        "linkx_homophily": ("homophily", "linkx_homophily"), "adjusted_homophily": ("homophily", "adjusted_homophily"),
        "edge_label_informativeness": ("label_informativeness", "edge_label_informativeness"),
        "node_label_informativeness": ("label_informativeness", "node_label_informativeness"),
        "register_func": ("_native.registry", "register_func"), "get_global_func": ("_native.registry", "get_global_func"),
        "list_global_func_names": ("_native.registry", "list_global_func_names"),
        "extract_ext_funcs": ("_native.registry", "extract_ext_funcs"), "TypeCode": ("_native.registry", "TypeCode"),
        "apply_each": ("utils.containers", "apply_each"), "copy_u": ("ops", "copy_u"), "copy_v": ("ops", "copy_v"),
        "set_src_lazy_features": ("dataloading.base", "set_src_lazy_features"),
        "set_dst_lazy_features": ("dataloading.base", "set_dst_lazy_features"),
        "set_node_lazy_features": ("dataloading.base", "set_node_lazy_features"),
        "set_edge_lazy_features": ("dataloading.base", "set_edge_lazy_features"),
        "load_backend": ("backend", "load_backend"), "backend_name": ("backend", "backend_name"),
        "use_libxsmm": ("global_config", "use_libxsmm"), "is_libxsmm_enabled": ("global_config", "is_libxsmm_enabled"),
    }
    if name in _late:
        modname, attr = _late[name]
        mod = importlib.import_module(f"{__name__}.{modname}")
        val = getattr(mod, attr)
        globals()[name] = val
        return val
    raise AttributeError(f"module '{__name__}' has no attribute '{name}'")
