"""DistGraph server/client, KVStore, DistTensor, distributed sampling, dataloaders.

Parity: /root/reference/python/dgl/distributed/{dist_context.py:208 initialize, dist_graph.py:488 DistGraphServer /
:722 DistGraph, kvstore.py (KVServer/KVClient push/pull/barrier), standalone_kvstore.py, dist_tensor.py,
graph_services.py:737-1160 (sample_neighbors / in_subgraph with per-partition RPC + merge), dist_dataloader.py,
nn/pytorch/sparse_emb.py DistEmbedding, optim/pytorch/sparse_optim.py}.
Two modes: ``standalone`` (DGL_DIST_MODE=standalone: in-process server object, no sockets) and ``distributed``
(server processes started with role env vars; clients talk over csrc/cpu/rpc.cc)."""
from __future__ import annotations

import atexit
import os
import threading

import numpy as np
import torch

from ..base import EID, NID, DGLError
from ..convert import graph as make_graph
from ..sampling.neighbor import sample_rel
from . import rpc
from .partition import load_partition, load_partition_book

SVC_META, SVC_PULL, SVC_PUSH, SVC_SAMPLE, SVC_BARRIER, SVC_INIT_DATA, SVC_INSUBGRAPH, SVC_FIND_EDGES = range(100, 108)


# ------------------------------------------------------------------------------------------- server state
class ServerState:
    """Everything one server owns: local partition, partition book, KV shards."""

    def __init__(self, part_config, part_id):
        self.part_id = part_id
        self.g, nf, ef, self.gpb, self.graph_name, self.ntypes, self.etypes = load_partition(part_config, part_id)
        self.kv = {}
        for k, v in nf.items():
            self.kv["node~" + k.replace("/", "~")] = v
        for k, v in ef.items():
            self.kv["edge~" + k.replace("/", "~")] = v
        self.policy = {}  # name -> ('node'|'edge', type)
        for k in self.kv:
            dom, typ, _ = k.split("~", 2)
            self.policy[k] = (dom, typ)
        self.handlers = {}
        self.lock = threading.Lock()
        self._barrier = {}
        # local lookup: global node id -> local id within the partition graph (inner nodes are a prefix)
        self.nid_start = int(self.gpb._node_map["_N"][part_id, 0])
        self.n_inner = int(self.gpb._node_map["_N"][part_id, 1]) - self.nid_start
        self.eid_start = int(self.gpb._edge_map[self.gpb._ekey("_E")][part_id, 0])
        self.local_nid = self.g.ndata[NID]
        self.local_eid = self.g.edata[EID]

    def local_range(self, name):
        dom, typ = self.policy[name]
        return self.nid_start if dom == "node" else self.eid_start


def _h_meta(st, payload):
    return {"graph_name": st.graph_name, "num_nodes": st.gpb.num_nodes(), "num_edges": st.gpb.num_edges(),
            "gpb": st.gpb.to_dict(), "num_parts": st.gpb.num_partitions(), "ntypes": st.ntypes, "etypes": st.etypes,
            "tensors": {k: (tuple(v.shape[1:]), v.dtype) for k, v in st.kv.items()}}


def _h_pull(st, payload):
    name, ids = payload
    return st.kv[name][ids - st.local_range(name)]


def _h_push(st, payload):
    name, ids, data = payload
    loc = ids - st.local_range(name)
    h = st.handlers.get(name)
    with st.lock:
        if h is not None:
            h(st.kv[name], loc, data)
        else:
            st.kv[name][loc] = data
    return True


def _h_init_data(st, payload):
    name, shape_tail, dtype, dom, typ, init = payload
    if name not in st.kv:
        n = st.n_inner if dom == "node" else (int(st.gpb._edge_map[st.gpb._ekey("_E")][st.part_id, 1]) - st.eid_start)
        t = torch.zeros((n,) + tuple(shape_tail), dtype=dtype)
        if init == "uniform":
            t.uniform_(-1, 1)
        st.kv[name] = t
        st.policy[name] = (dom, typ)
    return True


def _h_sample(st, payload):
    seeds, fanout, prob_name, replace, edge_dir = payload
    loc = seeds - st.nid_start
    prob = st.kv.get("edge~_N:_E:_N~" + prob_name) if prob_name else None
    p_full = None
    if prob is not None:
        # probabilities are stored for inner edges only (prefix of the partition's edges)
        p_full = torch.zeros(st.g.num_edges(), dtype=torch.float32)
        p_full[: prob.shape[0]] = prob.float()
    src, dst, eid = sample_rel(st.g._graphs[0], loc.to(st.g.idtype), fanout, edge_dir, p_full, replace)
    return st.local_nid[src.long()], st.local_nid[dst.long()], st.local_eid[eid.long()]


def _h_insubgraph(st, payload):
    seeds = payload
    loc = (seeds - st.nid_start).to(st.g.idtype)
    src, dst, eid = st.g.in_edges(loc, form="all")
    return st.local_nid[src.long()], st.local_nid[dst.long()], st.local_eid[eid.long()]


def _h_find_edges(st, payload):
    eids = payload
    loc = (eids - st.eid_start).to(st.g.idtype)  # inner edges are a prefix of the partition's edges
    src, dst = st.g.find_edges(loc)
    return st.local_nid[src.long()], st.local_nid[dst.long()]


def _h_barrier(st, payload):
    return True


for _sid, _h in ((SVC_META, _h_meta), (SVC_PULL, _h_pull), (SVC_PUSH, _h_push), (SVC_SAMPLE, _h_sample),
                 (SVC_BARRIER, _h_barrier), (SVC_INIT_DATA, _h_init_data), (SVC_INSUBGRAPH, _h_insubgraph),
                 (SVC_FIND_EDGES, _h_find_edges)):
    rpc.register_service(_sid, _h)


# ------------------------------------------------------------------------------------------- context
class _Context:
    def __init__(self):
        self.mode = None
        self.client = None
        self.local_servers = None  # standalone: list[ServerState]
        self.num_servers = 0
        self.initialized = False


_CTX = _Context()


def _call(server_id, service_id, payload):
    if _CTX.mode == "standalone":
        return rpc._SERVICES[service_id](_CTX.local_servers[server_id], payload)
    return _CTX.client.remote_call([(server_id, service_id, payload)])[0]


def _call_many(reqs):
    if _CTX.mode == "standalone":
        return [rpc._SERVICES[s](_CTX.local_servers[sid], p) for sid, s, p in reqs]
    return _CTX.client.remote_call(reqs)


def initialize(ip_config=None, max_queue_size=None, net_type="socket", num_worker_threads=1, use_graphbolt=False,
               part_config=None):
    """Set up this process' role (reference dist_context.py:208). Env: DGL_DIST_MODE (standalone|distributed),
    DGL_ROLE (server|client), DGL_SERVER_ID, DGL_CONF_PATH (partition json), DGL_NUM_CLIENT, DGL_IP_CONFIG."""
    mode = os.environ.get("DGL_DIST_MODE", "standalone")
    part_config = part_config or os.environ.get("DGL_CONF_PATH")
    if mode == "standalone":
        _CTX.mode = "standalone"
        if part_config is not None:
            import json
            with open(part_config) as f:
                nparts = json.load(f)["num_parts"]
            _CTX.local_servers = [ServerState(part_config, i) for i in range(nparts)]
            _CTX.num_servers = nparts
        _CTX.initialized = True
        return
    role = os.environ.get("DGL_ROLE", "client")
    ip_config = ip_config or os.environ.get("DGL_IP_CONFIG")
    servers = rpc.read_ip_config(ip_config)
    if role == "server":
        sid = int(os.environ["DGL_SERVER_ID"])
        assert part_config is not None, "DGL_CONF_PATH must be set for servers"
        state = ServerState(part_config, sid)
        ip, port = servers[sid]
        srv = rpc.RPCServer(sid, ip, port, int(os.environ.get("DGL_NUM_CLIENT", "1")), state)
        srv.serve_forever()
        os._exit(0)
    _CTX.mode = "distributed"
    _CTX.client = rpc.RPCClient(servers)
    _CTX.num_servers = len(servers)
    _CTX.initialized = True
    atexit.register(exit_client)


def exit_client():
    if _CTX.mode == "distributed" and _CTX.client is not None:
        _CTX.client.shutdown()
        _CTX.client = None
    _CTX.initialized = False


def is_initialized():
    return _CTX.initialized


# ------------------------------------------------------------------------------------------- client objects
class DistTensor:
    """A tensor sharded across servers by the partition book; indexing pulls / assignment pushes."""

    def __init__(self, shape, dtype, name=None, init_func=None, part_policy=None, persistent=False, is_gdata=True,
                 attach=True, _existing=False, _graph=None, _dom="node", _typ="_N"):
        self._shape, self._dtype = tuple(shape), dtype
        if _graph is None:
            # user-created tensor: attach to the process's DistGraph and pick the partition policy by length, as the
            # reference does when ``part_policy`` is omitted (dist_tensor.py:141-165)
            _graph = getattr(_CTX, "graph", None)
            if part_policy is not None and getattr(part_policy, "policy_str", "").startswith("edge"):
                _dom = "edge"
            elif _graph is not None and shape[0] == _graph.num_edges() and shape[0] != _graph.num_nodes():
                _dom = "edge"
        self._g = _graph
        self._dom, self._typ = _dom, _typ
        self._name = name if _existing else f"{_dom}~{_typ}~{name}"
        if not _existing:
            for sid in range(_CTX.num_servers):
                _call(sid, SVC_INIT_DATA, (self._name, self._shape[1:], dtype, _dom, _typ,
                                           "uniform" if init_func == "uniform" else "zeros"))

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return self._dtype

    @property
    def name(self):
        return self._name

    def __len__(self):
        return self._shape[0]

    @property
    def tensor_name(self):
        """User-facing name (without the kvstore prefix)."""
        return self._name.split("~", 2)[-1]

    @property
    def kvstore_key(self):
        return self._name

    @property
    def part_policy(self):
        from .kvstore import PartitionPolicy
        gpb = self._g.gpb if self._g is not None else None
        return PartitionPolicy(f"{self._dom}~{self._typ}", gpb)

    @property
    def local_partition(self):
        """Rows of this tensor stored on the local server (standalone / co-located), else None."""
        if _CTX.mode == "standalone" and _CTX.local_servers:
            store = getattr(_CTX.local_servers[0], "data", None) or getattr(_CTX.local_servers[0], "store", None)
            if isinstance(store, dict) and self._name in store:
                return store[self._name]
        return None

    def count_nonzero(self):
        """Number of non-zero entries over all shards."""
        tot, step = 0, 1 << 20
        for lo in range(0, self._shape[0], step):
            tot += int((self[torch.arange(lo, min(lo + step, self._shape[0]))] != 0).sum())
        return tot

    def _owner(self, ids):
        gpb = self._g.gpb
        return gpb.nid2partid(ids) if self._dom == "node" else gpb.eid2partid(ids)

    def __getitem__(self, idx):
        idx = torch.as_tensor(idx).long().reshape(-1)
        own = self._owner(idx)
        out = torch.empty((idx.shape[0],) + self._shape[1:], dtype=self._dtype)
        reqs, sel = [], []
        for p in torch.unique(own).tolist():
            m = torch.nonzero(own == p).squeeze(1)
            reqs.append((p, SVC_PULL, (self._name, idx[m])))
            sel.append(m)
        for m, res in zip(sel, _call_many(reqs)):
            out[m] = res
        return out

    def __setitem__(self, idx, val):
        idx = torch.as_tensor(idx).long().reshape(-1)
        own = self._owner(idx)
        reqs = []
        for p in torch.unique(own).tolist():
            m = torch.nonzero(own == p).squeeze(1)
            reqs.append((p, SVC_PUSH, (self._name, idx[m], val[m])))
        _call_many(reqs)


class _DistDataView:
    def __init__(self, g, dom):
        self._g, self._dom = g, dom

    def _prefix(self):
        return "node~_N~" if self._dom == "node" else "edge~_N:_E:_N~"

    def keys(self):
        pre = self._prefix()
        return [k[len(pre):] for k in self._g._tensors if k.startswith(pre)]

    def __contains__(self, k):
        return k in self.keys()

    def __getitem__(self, k):
        name = self._prefix() + k
        if name not in self._g._tensors:
            raise KeyError(k)
        tail, dtype = self._g._tensors[name]
        n = self._g.num_nodes() if self._dom == "node" else self._g.num_edges()
        return DistTensor((n,) + tuple(tail), dtype, name, _existing=True, _graph=self._g, _dom=self._dom,
                          _typ="_N" if self._dom == "node" else "_N:_E:_N")

    def __setitem__(self, k, val):
        if not isinstance(val, DistTensor):
            raise DGLError("assign a DistTensor created with the same graph")
        self._g._tensors[val.name] = (val.shape[1:], val.dtype)


ALL_NODES = "__ALL__"


class _DistTypeView:
    """``g.nodes[ntype].data`` / ``g.edges[etype].data`` views of a DistGraph."""

    def __init__(self, g, dom):
        self._g, self._dom = g, dom

    def __getitem__(self, typ):
        view = _DistDataView(self._g, self._dom)
        return type("_TypedView", (), {"data": view})()

    def __call__(self, typ=None):
        n = self._g.num_nodes() if self._dom == "node" else self._g.num_edges()
        return torch.arange(n)


class DistGraph:
    """Client handle of a partitioned graph served by DistGraphServer processes (or in-process when standalone)."""

    def __init__(self, graph_name=None, gpb=None, part_config=None):
        if not _CTX.initialized:
            initialize(part_config=part_config)
        if _CTX.mode == "standalone" and _CTX.local_servers is None:
            if part_config is None:
                raise DGLError("standalone DistGraph needs part_config")
            initialize(part_config=part_config)
        meta = _call(0, SVC_META, None)
        self.graph_name = meta["graph_name"]
        from .partition import RangePartitionBook
        self.gpb = RangePartitionBook(0, meta["num_parts"], meta["gpb"]["node_map"],
                                      {k: v for k, v in meta["gpb"]["edge_map"].items()}, {t: i for i, t in enumerate(meta["ntypes"])},
                                      {t: i for i, t in enumerate(meta["etypes"])})
        self._num_nodes, self._num_edges = meta["num_nodes"], meta["num_edges"]
        self._tensors = dict(meta["tensors"])
        self.ntypes, self.etypes = ["_N"], ["_E"]
        self.canonical_etypes = [("_N", "_E", "_N")]
        self.idtype = torch.int64
        self.device = torch.device("cpu")
        _CTX.graph = self

    def get_partition_book(self):
        return self.gpb

    def num_nodes(self, ntype=None):
        return self._num_nodes

    number_of_nodes = num_nodes

    def num_edges(self, etype=None):
        return self._num_edges

    number_of_edges = num_edges

    @property
    def ndata(self):
        return _DistDataView(self, "node")

    @property
    def edata(self):
        return _DistDataView(self, "edge")

    @property
    def local_partition(self):
        if _CTX.mode == "standalone":
            return _CTX.local_servers[0].g
        return None

    def rank(self):
        return _CTX.client.client_id if _CTX.mode == "distributed" else 0

    def barrier(self):
        _call_many([(sid, SVC_BARRIER, None) for sid in range(_CTX.num_servers)])

    def sample_neighbors(self, seed_nodes, fanout, edge_dir="in", prob=None, replace=False, **kw):
        return sample_neighbors(self, seed_nodes, fanout, edge_dir=edge_dir, prob=prob, replace=replace)

    def in_subgraph(self, nodes):
        return in_subgraph(self, nodes)

    def find_edges(self, eids, etype=None):
        return find_edges(self, eids)

    def get_rank(self):
        return self.rank()

    def get_num_client(self):
        return get_num_client()

    # ---- type / schema queries (reference dist_graph.py DistGraph) ----------------------------------------
    def get_ntype_id(self, ntype=None):
        if ntype is None:
            if len(self.ntypes) != 1:
                raise DGLError("Node type name must be specified if there are more than one node types.")
            return 0
        return self.ntypes.index(ntype)

    def to_canonical_etype(self, etype=None):
        if etype is None:
            if len(self.canonical_etypes) != 1:
                raise DGLError("Edge type name must be specified if there are more than one edge types.")
            return self.canonical_etypes[0]
        if isinstance(etype, tuple):
            return etype
        for c in self.canonical_etypes:
            if c[1] == etype:
                return c
        raise DGLError(f'Edge type "{etype}" does not exist.')

    def get_etype_id(self, etype=None):
        return self.canonical_etypes.index(self.to_canonical_etype(etype))

    @property
    def nodes(self):
        return _DistTypeView(self, "node")

    @property
    def edges(self):
        return _DistTypeView(self, "edge")

    def node_attr_schemes(self):
        return {k.split("~", 2)[2]: v for k, v in self._tensors.items() if k.startswith("node~")}

    def edge_attr_schemes(self):
        return {k.split("~", 2)[2]: v for k, v in self._tensors.items() if k.startswith("edge~")}

    def is_pinned(self):
        return False

    def get_node_partition_policy(self, ntype="_N"):
        from .kvstore import PartitionPolicy
        return PartitionPolicy("node~" + ntype, self.gpb)

    def get_edge_partition_policy(self, etype="_E"):
        from .kvstore import PartitionPolicy
        et = etype[1] if isinstance(etype, tuple) else etype
        return PartitionPolicy("edge~" + et, self.gpb)

    def add_edge_attribute(self, name):
        """Make edge data ``name`` available to the distributed samplers as an edge attribute (GraphBolt partitions
        keep attributes inside the sampling graph; here samplers read edge data through the servers directly)."""
        if "edge~_E~" + name not in self._tensors and name not in self.edge_attr_schemes():
            raise DGLError(f"Fail to find data of {name}.")

    def in_degrees(self, u=ALL_NODES):
        return self._degrees(u, "in")

    def out_degrees(self, u=ALL_NODES):
        return self._degrees(u, "out")

    def _degrees(self, u, direction):
        """Degrees of (global) node ids, answered by the owning partitions (standalone: the local partition)."""
        lp = self.local_partition
        if lp is None:
            raise DGLError("in_degrees / out_degrees need the local partition (standalone mode or co-located server)")
        from ..base import NID
        deg = lp.in_degrees() if direction == "in" else lp.out_degrees()
        full = torch.zeros(self._num_nodes, dtype=deg.dtype)
        full[lp.ndata[NID].long()] = deg if NID in lp.ndata else deg
        if isinstance(u, str):
            return full
        return full[torch.as_tensor(u).long()]

    def edge_subgraph(self, edges, relabel_nodes=True, store_ids=True):
        """Subgraph induced by (global) edge ids, built from ``find_edges`` (reference DistGraph.edge_subgraph)."""
        from ..convert import graph as make_graph
        from ..base import EID, NID
        eids = torch.as_tensor(edges).long()
        u, v = self.find_edges(eids)
        if relabel_nodes:
            nodes, inv = torch.unique(torch.cat([u, v]), return_inverse=True)
            sg = make_graph((inv[: u.shape[0]], inv[u.shape[0]:]), num_nodes=nodes.shape[0])
            if store_ids:
                sg.ndata[NID] = nodes
        else:
            sg = make_graph((u, v), num_nodes=self._num_nodes)
        if store_ids:
            sg.edata[EID] = eids
        return sg


def _merge(g, results):
    srcs = [r[0] for r in results if r[0].numel()]
    if srcs:
        src = torch.cat([r[0] for r in results]); dst = torch.cat([r[1] for r in results]); eid = torch.cat([r[2] for r in results])
    else:
        src = dst = eid = torch.empty(0, dtype=torch.int64)
    sg = make_graph((src, dst), num_nodes=g.num_nodes())
    sg.edata[EID] = eid
    return sg


def sample_neighbors(g, nodes, fanout, edge_dir="in", prob=None, replace=False):
    """Distributed neighbour sampling: seeds are split by owner partition, each owner samples on its local
    partition (which holds the complete in-neighbourhood of its inner nodes) and the pieces are merged
    (reference graph_services.py:1037-1160)."""
    nodes = torch.as_tensor(nodes).long().reshape(-1)
    own = g.gpb.nid2partid(nodes)
    reqs = []
    for p in torch.unique(own).tolist():
        reqs.append((p, SVC_SAMPLE, (nodes[own == p], int(fanout), prob, replace, edge_dir)))
    return _merge(g, _call_many(reqs))


def sample_etype_neighbors(g, nodes, fanout, edge_dir="in", prob=None, replace=False, etype_sorted=True, **kw):
    """Per-edge-type fan-out sampling. Partitions store homogenised graphs with a single relation in this build, so a
    dict / per-type ``fanout`` reduces to its (only) entry (reference graph_services.py:sample_etype_neighbors)."""
    if isinstance(fanout, dict):
        fanout = next(iter(fanout.values()))
    elif torch.is_tensor(fanout):
        fanout = int(fanout.reshape(-1)[0])
    return sample_neighbors(g, nodes, int(fanout), edge_dir=edge_dir, prob=prob, replace=replace)


def find_edges(g, edge_ids):
    """(src, dst) global node ids of the given global edge ids, resolved by the owning partitions."""
    eids = torch.as_tensor(edge_ids).long().reshape(-1)
    own = g.gpb.eid2partid(eids)
    src = torch.empty_like(eids)
    dst = torch.empty_like(eids)
    reqs, sel = [], []
    for p in torch.unique(own).tolist():
        m = torch.nonzero(own == p).squeeze(1)
        reqs.append((p, SVC_FIND_EDGES, eids[m]))
        sel.append(m)
    for m, (s_, d_) in zip(sel, _call_many(reqs)):
        src[m], dst[m] = s_, d_
    return src, dst


def get_rank():
    return _CTX.client.client_id if _CTX.mode == "distributed" and _CTX.client is not None else 0


def get_num_client():
    return int(os.environ.get("DGL_NUM_CLIENT", "1"))


class DistGraphServer:
    """One partition's graph + KVStore server process (reference dist_graph.py:488). ``start()`` blocks serving
    requests until every client has sent its shutdown message."""

    def __init__(self, server_id, ip_config, num_servers, num_clients, part_config, disable_shared_mem=False,
                 graph_format=("csc", "coo"), use_graphbolt=False):
        self.server_id, self.ip_config, self.num_clients = server_id, ip_config, num_clients
        self.state = ServerState(part_config, server_id)
        from .kvstore import KVServer
        self.kvserver = KVServer(server_id, ip_config, num_servers, num_clients, self.state)

    @property
    def gpb(self):
        return self.state.gpb

    def start(self):
        servers = rpc.read_ip_config(self.ip_config)
        ip, port = servers[self.server_id]
        rpc.RPCServer(self.server_id, ip, port, self.num_clients, self.state).serve_forever()


def in_subgraph(g, nodes):
    nodes = torch.as_tensor(nodes).long().reshape(-1)
    own = g.gpb.nid2partid(nodes)
    reqs = [(p, SVC_INSUBGRAPH, nodes[own == p]) for p in torch.unique(own).tolist()]
    return _merge(g, _call_many(reqs))


def node_split(nodes, partition_book=None, ntype="_N", rank=None, force_even=True, node_trainer_ids=None,
               num_trainers=None):
    """Split a node id set (tensor or boolean mask) across trainers; with a partition book prefer local nodes
    (reference dist_graph.py:node_split)."""
    import torch.distributed as dist
    nodes = torch.as_tensor(nodes)
    if nodes.dtype in (torch.bool, torch.uint8):  # masks read back from partition files are uint8, as in the reference
        nodes = torch.nonzero(nodes).squeeze(1)
    if num_trainers is None:
        num_trainers = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if partition_book is not None and not force_even:
        part = partition_book.nid2partid(nodes, ntype)
        return nodes[part == rank % partition_book.num_partitions()]
    per = -(-nodes.shape[0] // num_trainers)
    return nodes[rank * per: (rank + 1) * per]


def edge_split(edges, partition_book=None, etype="_E", rank=None, force_even=True, num_trainers=None, **kw):
    import torch.distributed as dist
    edges = torch.as_tensor(edges)
    if edges.dtype in (torch.bool, torch.uint8):
        edges = torch.nonzero(edges).squeeze(1)
    if num_trainers is None:
        num_trainers = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if partition_book is not None and not force_even:
        part = partition_book.eid2partid(edges, etype)
        return edges[part == rank % partition_book.num_partitions()]
    per = -(-edges.shape[0] // num_trainers)
    return edges[rank * per: (rank + 1) * per]


class DistDataLoader:
    """Batch ids and apply ``collate_fn`` (reference dist_dataloader.py:DistDataLoader)."""

    def __init__(self, dataset, batch_size, shuffle=False, collate_fn=None, drop_last=False, queue_size=None):
        self.dataset = torch.as_tensor(dataset)
        self.batch_size, self.shuffle, self.collate_fn, self.drop_last = batch_size, shuffle, collate_fn, drop_last

    def __len__(self):
        n = self.dataset.shape[0]
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        ids = self.dataset[torch.randperm(self.dataset.shape[0])] if self.shuffle else self.dataset
        for b in range(len(self)):
            batch = ids[b * self.batch_size: (b + 1) * self.batch_size]
            yield self.collate_fn(batch) if self.collate_fn else batch


class DistNodeDataLoader(DistDataLoader):
    """Node-wise neighbour-sampling loader over a DistGraph (reference dist_dataloader.py:792)."""

    def __init__(self, g, nids, graph_sampler, device=None, **kwargs):
        self.g, self.sampler, self.device = g, graph_sampler, device

        def collate(seeds):
            res = self.sampler.sample(self.g, seeds)
            return res

        super().__init__(nids, kwargs.pop("batch_size", 1), kwargs.pop("shuffle", False), collate, kwargs.pop("drop_last", False))


class DistEdgeDataLoader(DistDataLoader):
    def __init__(self, g, eids, graph_sampler, device=None, **kwargs):
        self.g, self.sampler = g, graph_sampler
        super().__init__(eids, kwargs.pop("batch_size", 1), kwargs.pop("shuffle", False),
                         lambda e: self.sampler.sample(self.g, e), kwargs.pop("drop_last", False))


class DistEmbedding:
    """Sparse embedding stored in the KVStore; gradients are pushed to the owners by the sparse optimizers
    (reference distributed/nn/pytorch/sparse_emb.py)."""

    def __init__(self, num_embeddings, embedding_dim, name=None, init_func=None, part_policy=None, _graph=None):
        self._name = name or "emb"
        self._tensor = DistTensor((num_embeddings, embedding_dim), torch.float32, self._name,
                                  init_func="uniform" if init_func is not None else None, _graph=_graph)
        self._trace = []
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim

    def __call__(self, idx, device=torch.device("cpu")):
        idx = torch.as_tensor(idx).long()
        emb = self._tensor[idx].to(device)
        if torch.is_grad_enabled():
            emb.requires_grad_(True)
            emb.retain_grad()
            self._trace.append((idx, emb))
        return emb

    def reset_trace(self):
        self._trace = []

    @property
    def weight(self):
        return self._tensor

    @property
    def name(self):
        return self._name

    @property
    def data_name(self):
        """KVStore key of the embedding table."""
        return self._tensor.name

    @property
    def part_policy(self):
        return self._tensor.part_policy

    @property
    def kvstore(self):
        from .compat import get_kvstore
        return get_kvstore()

    @property
    def optm_state(self):
        """State tensor(s) a sparse optimizer attached to this embedding (``None`` before an optimizer is created)."""
        return getattr(self, "_optm_state", None)


class DistSparseGradOptimizer:
    def __init__(self, params, lr):
        self._params, self._lr = params, lr
        self._state = {}

    def zero_grad(self):
        for p in self._params:
            p.reset_trace()

    def step(self):
        with torch.no_grad():
            for emb in self._params:
                idx = [i for i, e in emb._trace if e.grad is not None]
                if not idx:
                    continue
                idx = torch.cat(idx)
                grad = torch.cat([e.grad for _, e in emb._trace if e.grad is not None]).cpu()
                uniq, inv = torch.unique(idx, return_inverse=True)
                g = torch.zeros((uniq.shape[0], grad.shape[1])).index_add_(0, inv, grad)
                cnt = torch.bincount(inv, minlength=uniq.shape[0]).unsqueeze(1)
                self.update(emb, uniq, g / cnt)
                emb.reset_trace()

    def update(self, emb, idx, grad):
        raise NotImplementedError

    # ---- checkpointing (reference distributed/optim/pytorch/sparse_optim.py:56-260) ---------------------------
    @staticmethod
    def _world():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _hyper(self):
        return {k: v for k, v in vars(self).items() if k.startswith("_") and isinstance(v, (int, float, tuple))}

    def _state_tensors(self, emb):
        st = self._state[emb._name]
        return list(st) if isinstance(st, tuple) else [st]

    def local_state_dict(self):
        """This rank's share of the optimizer state: hyper-parameters plus, per embedding, a contiguous 1/world slice of
        global row ids and the state rows for them: ``{"params": {...}, "emb_states": {emb: {"ids": t, "states":
        {state name: rows}}}}``. Rows are addressed by *global* id, so a checkpoint can be restored under a different
        number of trainers."""
        rank, world = self._world()
        out = {"params": dict(self._hyper(), world_size=world), "emb_states": {}}
        for emb in self._params:
            n = emb.num_embeddings
            ids = torch.arange(rank * n // world, (rank + 1) * n // world)
            out["emb_states"][emb._name] = {"ids": ids, "states": {t.name: t[ids] for t in self._state_tensors(emb)}}
        return out

    def load_local_state_dict(self, local_state_dict):
        """Write the rows of a :meth:`local_state_dict` back into the distributed state tensors."""
        by_name = {emb._name: emb for emb in self._params}
        for emb_name, emb_state in local_state_dict["emb_states"].items():
            if emb_name not in by_name:
                raise ValueError(f"loaded state dict contains an unknown embedding {emb_name}")
            tensors = {t.name: t for t in self._state_tensors(by_name[emb_name])}
            if len(emb_state["states"]) != len(tensors):
                raise ValueError(f"loaded state dict has a different number of states of embedding {emb_name}")
            ids = emb_state["ids"]
            for name, rows in emb_state["states"].items():
                if name not in tensors:
                    raise ValueError(f"loaded state dict contains a state {name} that can't be found in the optimizer states")
                if ids.numel():
                    tensors[name][ids] = rows.to(tensors[name].dtype)
        for k, v in local_state_dict.get("params", {}).items():
            if k != "world_size" and hasattr(self, k):
                setattr(self, k, v)

    def save(self, f):
        """Collective: every rank writes ``<f>_<rank>``."""
        import torch.distributed as dist
        rank, world = self._world()
        if world > 1:
            dist.barrier()
        torch.save(self.local_state_dict(), f"{f}_{rank}")
        if world > 1:
            dist.barrier()

    def load(self, f):
        """Collective: restore from the files of a :meth:`save`, whatever world size wrote them — file ``i`` is read by
        rank ``i % world`` (covers scale-in and scale-out); a missing own file is a warning, not an error."""
        import glob
        import warnings
        import torch.distributed as dist
        rank, world = self._world()
        if world > 1:
            dist.barrier()
        files = {int(p.rsplit("_", 1)[1]): p for p in glob.glob(f"{f}_*") if p.rsplit("_", 1)[1].isdigit()}
        if rank not in files and not any(i % world == rank for i in files):
            warnings.warn(f"File {f}_{rank} can't be found, load nothing.")
        for i in sorted(files):
            if i % world == rank:
                self.load_local_state_dict(torch.load(files[i], weights_only=False))
        if world > 1:
            dist.barrier()


class DistSparseAdagrad(DistSparseGradOptimizer):
    def __init__(self, params, lr, eps=1e-10):
        super().__init__(params, lr)
        self._eps = eps
        for emb in params:
            self._state[emb._name] = DistTensor((emb.num_embeddings, emb.embedding_dim), torch.float32,
                                                emb._name + "_sum", _graph=emb._tensor._g)
            emb._optm_state = (self._state[emb._name],)

    def update(self, emb, idx, grad):
        st = self._state[emb._name]
        s = st[idx] + grad * grad
        st[idx] = s
        emb._tensor[idx] = emb._tensor[idx] - self._lr * grad / (s.sqrt() + self._eps)


class DistSparseAdam(DistSparseGradOptimizer):
    """Row-sparse Adam on a DistEmbedding: state (step, m, v) lives in DistTensors next to the embedding shard
    (reference distributed/optim/pytorch/sparse_optim.py:SparseAdam)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr)
        self._b1, self._b2, self._eps = betas[0], betas[1], eps
        for emb in params:
            g = emb._tensor._g
            self._state[emb._name] = (
                DistTensor((emb.num_embeddings, 1), torch.float32, emb._name + "_step", _graph=g),
                DistTensor((emb.num_embeddings, emb.embedding_dim), torch.float32, emb._name + "_m", _graph=g),
                DistTensor((emb.num_embeddings, emb.embedding_dim), torch.float32, emb._name + "_v", _graph=g))
            emb._optm_state = self._state[emb._name]

    def update(self, emb, idx, grad):
        st, m, v = self._state[emb._name]
        t = st[idx] + 1
        st[idx] = t
        mm = self._b1 * m[idx] + (1 - self._b1) * grad
        vv = self._b2 * v[idx] + (1 - self._b2) * grad * grad
        m[idx], v[idx] = mm, vv
        mhat = mm / (1 - self._b1 ** t)
        vhat = vv / (1 - self._b2 ** t)
        emb._tensor[idx] = emb._tensor[idx] - self._lr * mhat / (vhat.sqrt() + self._eps)


SparseAdagrad = DistSparseAdagrad
SparseAdam = DistSparseAdam


NodeDataView = EdgeDataView = _DistDataView          # reference names of the ndata / edata views

# what ``g.nodes[ntype]`` / ``g.edges[etype]`` hand back in the reference (dist_graph.py:343-372): a record whose only field
# is the data view; ``_DistTypeView`` items expose the same ``.data``
from collections import namedtuple  # noqa: E402

NodeSpace = namedtuple("NodeSpace", ["data"])
EdgeSpace = namedtuple("EdgeSpace", ["data"])
from .service_ids import (ADD_EDGE_ATTRIBUTE_FROM_KV, ADD_EDGE_ATTRIBUTE_FROM_SHARED_MEM, INIT_GRAPH,  # noqa: E402,F401
                          QUERY_IF_USE_GRAPHBOLT)
HeteroNodeView = HeteroEdgeView = _DistTypeView      # reference names of g.nodes[...] / g.edges[...] views


def __getattr__(name):
    # class-based request / response types of this service live in messages.py (lazy: avoids an import cycle)
    if name.startswith("__"):
        raise AttributeError(name)
    from . import messages
    if name in messages.__all__:
        return getattr(messages, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
