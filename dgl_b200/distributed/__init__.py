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
from .partition import dgl_partition_to_graphbolt, gb_convert_single_dgl_partition  # noqa: F401,E402
from .rpc import connect_to_server, start_server  # noqa: F401,E402


class NodeCollator:
    """Legacy collator kept for API parity (reference distributed/dist_dataloader.py NodeCollator): turns a list of seed
    nodes into (input_nodes, output_nodes, blocks) with a block sampler."""

    def __init__(self, g, nids, graph_sampler):
        self.g, self.nids, self.graph_sampler = g, nids, graph_sampler

    @property
    def dataset(self):
        return self.nids

    def collate(self, items):
        import torch
        items = torch.as_tensor(items) if not isinstance(items, dict) else items
        return self.graph_sampler.sample(self.g, items)


class EdgeCollator(NodeCollator):
    """Legacy edge collator (reference EdgeCollator): seeds are edge ids; the sampler must be an edge-prediction sampler."""

    def __init__(self, g, eids, graph_sampler, g_sampling=None, exclude=None, reverse_eids=None, reverse_etypes=None,
                 negative_sampler=None):
        super().__init__(g, eids, graph_sampler)
        from ..dataloading import as_edge_prediction_sampler
        if not hasattr(graph_sampler, "sample_blocks_for_edges") and (negative_sampler is not None or exclude is not None):
            self.graph_sampler = as_edge_prediction_sampler(graph_sampler, exclude=exclude, reverse_eids=reverse_eids,
                                                            reverse_etypes=reverse_etypes, negative_sampler=negative_sampler)

from . import compat, rpc_api  # noqa: F401,E402
from .compat import *  # noqa: F401,F403,E402
from .rpc_api import *  # noqa: F401,F403,E402  (reference distributed/__init__.py:`from .rpc import *`)
from .rpc_api import (client_barrier, connect_receiver, connect_receiver_finalize, copy_data_to_shared_memory,  # noqa: F401,E402
                      create_receiver, create_sender, fast_pull, finalize_receiver, finalize_sender, finalize_server,
                      recv_request, recv_response, recv_responses, recv_rpc_message, remote_call, remote_call_to_machine,
                      send_request, send_request_to_machine, send_requests_to_machine, send_response, send_rpc_message,
                      wait_for_senders)
from .rpc import read_ip_config  # noqa: F401,E402
# the context-aware variants stay the package-level names (the RPC-state accessors remain reachable as rpc_api.*)
from .dist_graph import get_num_client, get_rank  # noqa: F401,E402


_COMPAT_LAZY = ("SubgraphResponse", "FindEdgeResponse", "InDegreeResponse", "OutDegreeResponse", "SamplingRequest",
                "SamplingRequestEtype", "InSubgraphRequest", "EdgesRequest", "InDegreeRequest", "OutDegreeRequest")


def __getattr__(name):
    if name in _COMPAT_LAZY:
        return getattr(compat, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

# reference distributed/constants.py
MAX_QUEUE_SIZE = 20 * 1024 * 1024 * 1024
SERVER_EXIT = "server_exit"
DEFAULT_NTYPE = "_N"
DEFAULT_ETYPE = (DEFAULT_NTYPE, "_E", DEFAULT_NTYPE)
DGL2GB_EID = "_dgl2gb_eid"
GB_DST_ID = "_gb_dst_id"
