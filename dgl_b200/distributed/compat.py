"""Reference-named pieces of ``dgl.distributed`` that sit around the core classes (parity:
/root/reference/python/dgl/distributed/{id_map,graph_partition_book,kvstore,dist_context,graph_services,rpc_client,
partition}.py). Storage, request handlers and the sampling services live in ``dist_graph.py`` / ``kvstore.py``."""
from __future__ import annotations

import os
import socket

import numpy as np
import torch

from ..base import DGLError
from .kvstore import EDGE_PART_POLICY, NODE_PART_POLICY, KVClient, PartitionPolicy
from .kvstore import close_kvstore, default_pull_handler, default_push_handler, get_kvstore, init_kvstore  # noqa: F401,E402
from .partition import (cast_various_to_minimum_dtype_gb, init_type_per_edge, is_homogeneous,  # noqa: F401,E402
                        process_partitions)

__all__ = ["IdMap", "HeteroDataName", "parse_hetero_data_name", "NodePartitionPolicy", "EdgePartitionPolicy",
           "get_shared_mem_partition_book", "get_node_partition_from_book", "default_pull_handler",
           "default_push_handler", "init_kvstore", "get_kvstore", "close_kvstore", "local_ip4_addr_list",
           "get_local_machine_id", "get_local_usable_addr", "in_degrees", "out_degrees", "merge_graphs",
           "is_homogeneous", "init_type_per_edge", "cast_various_to_minimum_dtype_gb", "process_partitions",
           "set_initialized", "init_process", "finalize_client", "finalize_worker", "join_finalize_worker",
           "get_sampler_pool", "CustomPool", "MpCommand", "POLICY_DELIMITER", "Collator"]

POLICY_DELIMITER = "~"


# ---- homogeneous id <-> (type id, per-type id) ------------------------------------------------------------------
class IdMap:
    """Decode *homogeneous* ids of a partitioned heterograph. Every (partition, type) pair owns one contiguous id range
    (``id_ranges[type][p] = [start, end)``); an id maps to the type that owns its range and to its rank among the ids
    of that type over all partitions (reference id_map.py). Lookup is one ``searchsorted`` over the range starts."""

    def __init__(self, id_ranges):
        vals = [np.asarray(v) for v in id_ranges.values()]
        self.num_types = len(vals)
        if self.num_types == 0:
            return
        self.num_parts = vals[0].shape[0]
        order = sorted(range(self.num_types), key=lambda i: vals[i][0, 0])      # type ids follow the id layout
        starts, ends, tids, base = [], [], [], []
        for rank, i in enumerate(order):
            r = torch.as_tensor(vals[i]).long()
            sizes = r[:, 1] - r[:, 0]
            before = torch.cumsum(sizes, 0) - sizes                                # ids of this type in earlier parts
            starts.append(r[:, 0]); ends.append(r[:, 1])
            tids.append(torch.full((self.num_parts,), rank)); base.append(before)
        starts, ends, tids, base = (torch.cat(x) for x in (starts, ends, tids, base))
        o = torch.argsort(starts)
        self._start, self._end, self._tid, self._base = starts[o], ends[o], tids[o], base[o]
        if bool((self._start[1:] < self._end[:-1]).any()):
            raise DGLError("id ranges overlap")

    @property
    def torch_dtype(self):
        return torch.int64

    def __call__(self, ids):
        ids = torch.as_tensor(ids)
        if self.num_types == 0:
            return torch.zeros(ids.shape[0], dtype=ids.dtype), ids
        if ids.numel() == 0:
            return ids, ids
        slot = torch.searchsorted(self._start, ids.long(), right=True) - 1
        if bool((slot < 0).any()) or bool((ids.long() >= self._end[slot.clamp(min=0)]).any()):
            raise DGLError("some ids fall outside every (partition, type) range")
        return self._tid[slot].to(ids.dtype), (ids.long() - self._start[slot] + self._base[slot]).to(ids.dtype)


# ---- data names / policies ------------------------------------------------------------------------------------------
def _etype_str(t):
    return ":".join(t) if isinstance(t, tuple) else t


class HeteroDataName:
    """KVStore key of one distributed tensor: domain ('node' | 'edge') ~ entity type ~ name."""

    def __init__(self, is_node, entity_type, data_name):
        if not is_node and not (isinstance(entity_type, tuple) and len(entity_type) == 3):
            raise AssertionError(f"Expect canonical edge type in a triplet of string, but got {entity_type}.")
        self._policy = NODE_PART_POLICY if is_node else EDGE_PART_POLICY
        self._entity_type, self.data_name = entity_type, data_name

    @property
    def policy_str(self):
        return self._policy + POLICY_DELIMITER + _etype_str(self._entity_type)

    def is_node(self):
        return self._policy == NODE_PART_POLICY

    def is_edge(self):
        return self._policy == EDGE_PART_POLICY

    def get_type(self):
        return self._entity_type

    def get_name(self):
        return self.data_name

    def __str__(self):
        return self.policy_str + POLICY_DELIMITER + self.data_name


def parse_hetero_data_name(name):
    parts = name.split(POLICY_DELIMITER)
    assert len(parts) == 3 and parts[0] in (NODE_PART_POLICY, EDGE_PART_POLICY), \
        f"{name} is not a valid heterograph data name"
    is_node = parts[0] == NODE_PART_POLICY
    etype = parts[1] if is_node else tuple(parts[1].split(":"))
    return HeteroDataName(is_node, etype, parts[2])


class NodePartitionPolicy(PartitionPolicy):
    """Partition policy of one node type."""

    def __init__(self, partition_book, ntype="_N"):
        super().__init__(NODE_PART_POLICY + ":" + ntype, partition_book)


class EdgePartitionPolicy(PartitionPolicy):
    """Partition policy of one (canonical) edge type."""

    def __init__(self, partition_book, etype=("_N", "_E", "_N")):
        super().__init__(EDGE_PART_POLICY + ":" + _etype_str(etype), partition_book)


def get_shared_mem_partition_book(graph_name):
    """Partition book that a server published to shared memory under ``graph_name`` (``None`` if absent)."""
    from ..utils.shared_mem import shm_exists, tensors_from_shared_memory
    key = f"{graph_name}_gpb"
    if not shm_exists(f"/{key}_index"):
        return None
    from .partition import RangePartitionBook
    tensors, meta = tensors_from_shared_memory(key)
    node_map = {k.split("~", 1)[1]: v for k, v in tensors.items() if k.startswith("node~")}
    edge_map = {k.split("~", 1)[1]: v for k, v in tensors.items() if k.startswith("edge~")}
    return RangePartitionBook(meta["part_id"], meta["num_parts"], node_map, edge_map, meta["ntypes"], meta["etypes"])


def get_node_partition_from_book(book, device):
    """Range :class:`~dgl_b200.partition.NDArrayPartition` over all nodes with one part per graph partition."""
    from ..partition import NDArrayPartition
    ends = torch.stack([m[:, 1] for m in book._node_map.values()]).max(0)[0]
    ranges = torch.cat([torch.zeros(1, dtype=torch.long), ends]).to(device)
    return NDArrayPartition(int(ends[-1]), book.num_partitions(), mode="range", part_ranges=ranges)


# ---- kvstore module-level API ------------------------------------------------------------------------------------------


# ---- rpc_client helpers ------------------------------------------------------------------------------------------------
def local_ip4_addr_list():
    """IPv4 addresses of this host (loopback included)."""
    addrs = {"127.0.0.1"}
    try:
        for info in socket.getaddrinfo(socket.gethostname(), None, socket.AF_INET):
            addrs.add(info[4][0])
    except socket.gaierror:
        pass
    try:
        with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s:     # no packet is sent by connect() on UDP
            s.connect(("10.255.255.255", 1))
            addrs.add(s.getsockname()[0])
    except OSError:
        pass
    return addrs


def get_local_machine_id(server_namebook):
    """Id of the machine in ``{server id: [machine id, ip, port, ...]}`` whose ip is one of ours."""
    mine = local_ip4_addr_list()
    for entry in server_namebook.values():
        if entry[1] in mine:
            return entry[0]
    raise DGLError("this host does not appear in the ip config")


def get_local_usable_addr(probe_addr):
    """'ip:port' of a free local port reachable from ``probe_addr``."""
    ip = "127.0.0.1"
    try:
        with socket.socket(socket.AF_INET, socket.SOCK_DGRAM) as s:
            s.connect((probe_addr, 14))
            ip = s.getsockname()[0]
    except OSError:
        pass
    with socket.socket() as s:
        s.bind(("", 0))
        port = s.getsockname()[1]
    return f"{ip}:{port}"


# ---- graph_services functions -----------------------------------------------------------------------------------------
def in_degrees(g, v):
    """In-degrees of global node ids on a :class:`DistGraph`."""
    return g.in_degrees(v)


def out_degrees(g, u):
    return g.out_degrees(u)


def merge_graphs(res, num_nodes, exclude_edges=None):
    """Union of per-partition sampling results ``[(src, dst, eid), ...]`` as one graph over ``num_nodes`` nodes."""
    from ..base import EID
    from ..convert import graph as make_graph
    src = torch.cat([r[0] for r in res]) if res else torch.empty(0, dtype=torch.long)
    dst = torch.cat([r[1] for r in res]) if res else torch.empty(0, dtype=torch.long)
    eid = torch.cat([r[2] for r in res]) if res else torch.empty(0, dtype=torch.long)
    if exclude_edges is not None and eid.numel():
        keep = ~torch.isin(eid, torch.as_tensor(exclude_edges))
        src, dst, eid = src[keep], dst[keep], eid[keep]
    g = make_graph((src, dst), num_nodes=int(num_nodes))
    g.edata[EID] = eid
    return g


# ---- request / response classes of graph_services.py and dist_dataloader.Collator -----------------------------------
def _service_classes():
    """The sampling / degree / subgraph services are function handlers in ``dist_graph`` (``_h_sample`` ...). These
    classes give them the reference's Request/Response shape on top of the same handlers."""
    from . import dist_graph as dg
    from .rpc_api import Request, Response

    class SubgraphResponse(Response):
        def __init__(self, global_src, global_dst, global_eids):
            self.global_src, self.global_dst, self.global_eids = global_src, global_dst, global_eids

    class FindEdgeResponse(Response):
        def __init__(self, global_src, global_dst, order_id):
            self.global_src, self.global_dst, self.order_id = global_src, global_dst, order_id

    class InDegreeResponse(Response):
        def __init__(self, val, order_id):
            self.val, self.order_id = val, order_id

    class OutDegreeResponse(InDegreeResponse):
        pass

    def _triple(res):
        return SubgraphResponse(*res[:3]) if isinstance(res, (tuple, list)) else res

    class SamplingRequest(Request):
        def __init__(self, nodes, fan_out, edge_dir="in", prob=None, replace=False, use_graphbolt=False):
            self.seed_nodes, self.fan_out, self.edge_dir, self.prob, self.replace = nodes, fan_out, edge_dir, prob, replace

        def process_request(self, server_state):
            return _triple(dg._h_sample(server_state, (self.seed_nodes, self.fan_out, self.prob, self.replace,
                                                       self.edge_dir)))

    class SamplingRequestEtype(SamplingRequest):
        def __init__(self, nodes, fan_out, edge_dir="in", prob=None, replace=False, etype_sorted=True,
                     use_graphbolt=False):
            super().__init__(nodes, fan_out, edge_dir, prob, replace)
            self.etype_sorted = etype_sorted

    class InSubgraphRequest(Request):
        def __init__(self, nodes):
            self.seed_nodes = nodes

        def process_request(self, server_state):
            return _triple(dg._h_insubgraph(server_state, self.seed_nodes))

    class EdgesRequest(Request):
        def __init__(self, edge_ids, order_id):
            self.edge_ids, self.order_id = edge_ids, order_id

        def process_request(self, server_state):
            res = dg._h_find_edges(server_state, self.edge_ids)
            return FindEdgeResponse(res[0], res[1], self.order_id)

    class InDegreeRequest(Request):
        def __init__(self, n, order_id):
            self.n, self.order_id = n, order_id

        def _degree(self, server_state, direction):
            g = server_state.graph
            local = server_state.gpb.nid2localnid(self.n, server_state.gpb.partid) if hasattr(server_state, "gpb") else self.n
            return (g.in_degrees if direction == "in" else g.out_degrees)(local)

        def process_request(self, server_state):
            return InDegreeResponse(self._degree(server_state, "in"), self.order_id)

    class OutDegreeRequest(InDegreeRequest):
        def process_request(self, server_state):
            return OutDegreeResponse(self._degree(server_state, "out"), self.order_id)

    from .rpc_api import register_service
    for sid, req, res in ((SAMPLING_SERVICE_ID, SamplingRequest, SubgraphResponse),
                          (INSUBGRAPH_SERVICE_ID, InSubgraphRequest, SubgraphResponse),
                          (EDGES_SERVICE_ID, EdgesRequest, FindEdgeResponse),
                          (OUTDEGREE_SERVICE_ID, OutDegreeRequest, OutDegreeResponse),
                          (INDEGREE_SERVICE_ID, InDegreeRequest, InDegreeResponse),
                          (ETYPE_SAMPLING_SERVICE_ID, SamplingRequestEtype, SubgraphResponse)):
        register_service(sid, req, res)
    return {c.__name__: c for c in (SubgraphResponse, FindEdgeResponse, InDegreeResponse, OutDegreeResponse,
                                    SamplingRequest, SamplingRequestEtype, InSubgraphRequest, EdgesRequest,
                                    InDegreeRequest, OutDegreeRequest)}


from .service_ids import (EDGES_SERVICE_ID, ETYPE_SAMPLING_SERVICE_ID, INDEGREE_SERVICE_ID,  # noqa: E402,F401
                          INSUBGRAPH_SERVICE_ID, OUTDEGREE_SERVICE_ID, SAMPLING_SERVICE_ID)


class LocalSampledGraph:
    """What a partition-local sampling call yields (graph_services.py:LocalSampledGraph): global source / destination
    / edge ids of the sampled edges and, for typed sampling, their edge type ids."""

    def __init__(self, global_src, global_dst, global_eids, etype_ids=None):
        self.global_src, self.global_dst, self.global_eids, self.etype_ids = global_src, global_dst, global_eids, etype_ids


# per-process counters / slots the reference keeps at module level
DIST_TENSOR_ID = 0           # dist_tensor.py: suffix of anonymous DistTensor names
DATALOADER_ID = 0            # dist_dataloader.py: id of the next DistDataLoader in this process
SAMPLER_POOL = None          # dist_context.py: the sampler worker pool (none: samplers run in the trainer here)
NUM_SAMPLER_WORKERS = 0
INITIALIZED = False          # dist_context.py: flipped by initialize()
CANONICAL_ETYPE_DELIMITER = ":"                                            # graph_partition_book.py
EMB_STATES, WORLD_SIZE, IDS, PARAMS, STATES = "emb_states", "world_size", "ids", "params", "states"   # sparse_optim.py


def _dtype_tables():
    import torch
    from ..backend import data_type_dict
    names = {v: k for k, v in data_type_dict.items()}                                    # shared_mem_utils.DTYPE_DICT
    check = {"node_map": torch.int64, "edge_map": torch.int64, "inner_node": torch.uint8, "inner_edge": torch.uint8,
             "part_id": torch.int64, "_ID": torch.int64, "orig_id": torch.int64}         # partition.DTYPES_TO_CHECK
    return names, check


DTYPE_DICT, DTYPES_TO_CHECK = _dtype_tables()


_SERVICE_CLASSES = None


def __getattr__(name):
    global _SERVICE_CLASSES
    if name in ("SubgraphResponse", "FindEdgeResponse", "InDegreeResponse", "OutDegreeResponse", "SamplingRequest",
                "SamplingRequestEtype", "InSubgraphRequest", "EdgesRequest", "InDegreeRequest", "OutDegreeRequest"):
        if _SERVICE_CLASSES is None:
            _SERVICE_CLASSES = _service_classes()
        return _SERVICE_CLASSES[name]
    raise AttributeError(name)


class Collator:
    """Abstract collator of the pre-0.8 distributed loaders (reference dist_dataloader.py:Collator): ``dataset`` is
    what a loader iterates, ``collate(items)`` turns a batch of items into sampler output."""

    @property
    def dataset(self):
        raise NotImplementedError

    def collate(self, items):
        raise NotImplementedError


# ---- partition.py helpers ------------------------------------------------------------------------------------------------


# ---- dist_context -------------------------------------------------------------------------------------------------------
class MpCommand:
    """Commands understood by sampler worker processes of the reference's CustomPool."""
    INIT_RPC = 0
    SET_COLLATE_FN = 1
    CALL_BARRIER = 2
    DELETE_COLLATE_FN = 3
    CALL_COLLATE_FN = 4
    CALL_FN_ALL_WORKERS = 5
    FINALIZE_POOL = 6


class CustomPool:
    """The reference forks ``num_workers`` sampler processes per trainer and feeds them commands through queues.
    Distributed sampling here is served by the partition servers (native kernels), so the pool degenerates to calling
    the collate function inline; the class keeps the control surface (``set_collate_fn`` / ``submit_task`` /
    ``get_result`` / ``close`` / ``join``)."""

    def __init__(self, num_workers=0, rpc_config=None):
        self.num_workers = num_workers
        self._fns, self._results = {}, {}

    def set_collate_fn(self, func, dataloader_name):
        self._fns[dataloader_name] = func

    def submit_task(self, dataloader_name, args):
        self._results.setdefault(dataloader_name, []).append(self._fns[dataloader_name](*args))

    def submit_task_to_all_workers(self, func, args):
        func(*args)

    def get_result(self, dataloader_name, timeout=1800):
        return self._results[dataloader_name].pop(0)

    def delete_collate_fn(self, dataloader_name):
        self._fns.pop(dataloader_name, None)
        self._results.pop(dataloader_name, None)

    def call_barrier(self):
        return None

    def close(self):
        self._fns.clear()

    def join(self):
        return None


_SAMPLER_POOL = [None, 0]


def set_initialized(value=True):
    from . import dist_graph
    dist_graph._CTX.initialized = bool(value)


def init_process(rpc_config, mp_contexts):
    """Entry point of a sampler worker in the reference; connects the worker to the servers."""
    from . import dist_graph
    if not dist_graph._CTX.initialized:
        dist_graph.initialize(*rpc_config[:1])


def get_sampler_pool():
    """(pool, number of sampler workers) of this trainer."""
    if _SAMPLER_POOL[0] is None:
        _SAMPLER_POOL[0] = CustomPool(int(os.environ.get("DGL_DIST_NUM_SAMPLER", "0")))
        _SAMPLER_POOL[1] = _SAMPLER_POOL[0].num_workers
    return _SAMPLER_POOL[0], _SAMPLER_POOL[1]


def finalize_client():
    from . import dist_graph
    close_kvstore()
    dist_graph.exit_client()


def finalize_worker():
    pool = _SAMPLER_POOL[0]
    if pool is not None:
        pool.close()


def join_finalize_worker():
    pool = _SAMPLER_POOL[0]
    if pool is not None:
        pool.join()
        _SAMPLER_POOL[0] = None
