"""KVStore facade: named, partition-sharded tensors with pull / push / barrier.

Parity: /root/reference/python/dgl/distributed/kvstore.py (KVServer :806, KVClient :1016, PartitionPolicy in
graph_partition_book.py:1013-1110) and standalone_kvstore.py. The storage and the request handlers live in
``dist_graph.ServerState`` (one per partition); this module gives them the reference's class names and call shapes so
user code written against ``dgl.distributed.kvstore`` keeps working."""
from __future__ import annotations

import torch

from ..base import DGLError

from .service_ids import (BARRIER, BARRIER_MSG, COUNT_LOCAL_NONZERO, DELETE_DATA, DELETE_MSG, GET_PART_SHAPE,  # noqa: F401
                          GET_SHARED, GET_SHARED_MSG, INIT_DATA, INIT_MSG, KVSTORE_PULL, KVSTORE_PUSH, REGISTER_PULL,
                          REGISTER_PULL_MSG, REGISTER_PUSH, REGISTER_PUSH_MSG, SEND_META_TO_BACKUP,
                          SEND_META_TO_BACKUP_MSG)

KVCLIENT = None        # the reference's module-level client slot; this package keeps it in ``_CLIENT`` (get_kvstore())
NODE_PART_POLICY = "node"
EDGE_PART_POLICY = "edge"


class PartitionPolicy:
    """Maps global ids of one node/edge type to (partition, local id) through the partition book."""

    def __init__(self, policy_str, partition_book):
        if ":" in policy_str:
            dom, typ = policy_str.split(":", 1)
        else:
            dom, typ = policy_str, ("_N" if policy_str == NODE_PART_POLICY else "_N:_E:_N")
        if dom not in (NODE_PART_POLICY, EDGE_PART_POLICY):
            raise DGLError(f"unknown partition policy {policy_str!r}")
        self._dom, self._typ, self._gpb = dom, typ, partition_book
        self._policy_str = f"{dom}:{typ}"

    @property
    def policy_str(self):
        return self._policy_str

    @property
    def part_id(self):
        return self._gpb.partid

    @property
    def partition_book(self):
        return self._gpb

    @property
    def type_name(self):
        """Node type name, or the canonical edge type as a tuple."""
        return self._typ if self._dom == NODE_PART_POLICY else tuple(self._typ.split(":"))

    def is_node(self):
        return self._dom == NODE_PART_POLICY

    def get_data_name(self, name):
        """KVStore key of tensor ``name`` under this policy (a ``HeteroDataName``)."""
        from .compat import HeteroDataName
        return HeteroDataName(self.is_node(), self.type_name, name)

    def to_partid(self, id_tensor):
        if self._dom == NODE_PART_POLICY:
            return self._gpb.nid2partid(id_tensor, self._typ)
        return self._gpb.eid2partid(id_tensor, self._typ)

    def to_local(self, id_tensor):
        if self._dom == NODE_PART_POLICY:
            return self._gpb.nid2localnid(id_tensor, self._gpb.partid, self._typ)
        return self._gpb.eid2localeid(id_tensor, self._gpb.partid, self._typ)

    def _range_table(self):
        """[num_parts, 2] id ranges of THIS policy's node / edge type."""
        if self._dom == NODE_PART_POLICY:
            return self._gpb._node_map[self._typ]
        return self._gpb._edge_map[self._gpb._ekey(self._typ)]

    def get_part_size(self):
        m = self._range_table()
        return int(m[self._gpb.partid, 1] - m[self._gpb.partid, 0])

    def get_size(self):
        m = self._range_table()
        return int((m[:, 1] - m[:, 0]).sum())


class KVServer:
    """Server-side view of one partition's shard store (wraps a ``ServerState``)."""

    def __init__(self, server_id, ip_config=None, num_servers=1, num_clients=1, state=None):
        self.server_id, self.num_servers, self.num_clients = server_id, num_servers, num_clients
        self._state = state
        self.part_policy = {}

    def attach(self, state):
        self._state = state
        return self

    @property
    def data_store(self):
        return self._state.kv

    def is_backup_server(self):
        return False

    @property
    def part_id(self):
        return self._state.part_id if self._state is not None else self.server_id

    @property
    def barrier_count(self):
        return self._barrier_count if hasattr(self, "_barrier_count") else {}

    @barrier_count.setter
    def barrier_count(self, value):
        self._barrier_count = value

    @property
    def orig_data(self):
        """Names of tensors created by this server (all of them: there are no backup servers sharing them)."""
        return set(self._state.kv) if self._state is not None else set()

    @property
    def push_handlers(self):
        return {k: v for k, v in self._state.handlers.items() if k != "__pull__"} if self._state is not None else {}

    @property
    def pull_handlers(self):
        return dict(self._state.handlers.get("__pull__", {})) if self._state is not None else {}

    def count_local_nonzero(self, name):
        return int(torch.count_nonzero(self._state.kv[name]))

    def add_part_policy(self, policy):
        self.part_policy[policy.policy_str] = policy

    def init_data(self, name, policy_str, data_tensor=None):
        dom, typ = policy_str.split(":", 1) if ":" in policy_str else (policy_str, "_N")
        if data_tensor is not None:
            self._state.kv[name] = data_tensor
        self._state.policy[name] = (dom, typ)

    def register_push_handler(self, name, func):
        """``func(target_tensor, local_ids, data)`` replaces plain assignment for pushes to ``name``."""
        self._state.handlers[name] = func

    def find_policy(self, policy_str):
        return self.part_policy[policy_str]


class KVClient:
    """Client-side pull/push on named distributed tensors (works in standalone and distributed mode)."""

    def __init__(self, ip_config=None, num_servers=1, role="default"):
        from . import dist_graph as dg
        if not dg._CTX.initialized:
            raise DGLError("call dgl.distributed.initialize() first")
        self._dg = dg
        self._gpb = None
        self._meta = {}

    def map_shared_data(self, partition_book):
        self._gpb = partition_book
        meta = self._dg._call(0, self._dg.SVC_META, None)
        self._meta = dict(meta["tensors"])

    def _owner(self, name, ids):
        dom = name.split("~", 1)[0]
        return self._gpb.nid2partid(ids) if dom == "node" else self._gpb.eid2partid(ids)

    @property
    def num_servers(self):
        return self._dg._CTX.num_servers

    @property
    def group_count(self):
        """Servers per machine (one per partition here)."""
        return 1

    @property
    def data_store(self):
        """Local shard store in standalone mode (``{}`` when the shards live in server processes)."""
        st = getattr(self._dg._CTX, "local_state", None)
        return st.kv if st is not None else {}

    def add_data(self, name, tensor, part_policy=None):
        """Standalone-mode helper of the reference's SA KVClient: register an existing full tensor under ``name``."""
        full = self.init_data(name, tuple(tensor.shape), tensor.dtype, part_policy)
        self.push(full, torch.arange(tensor.shape[0]), tensor)
        return full

    def data_name_list(self):
        return list(self._meta)

    def get_data_meta(self, name):
        tail, dtype = self._meta[name]
        return dtype, tuple(tail), name.split("~", 1)[0]

    def init_data(self, name, shape, dtype, part_policy=None, init_func=None, is_gdata=True):
        dom = "node" if part_policy is None or part_policy.policy_str.startswith("node") else "edge"
        typ = "_N" if dom == "node" else "_N:_E:_N"
        full = name if "~" in name else f"{dom}~{typ}~{name}"
        for sid in range(self.num_servers):
            self._dg._call(sid, self._dg.SVC_INIT_DATA, (full, tuple(shape[1:]), dtype, dom, typ,
                                                        "uniform" if init_func == "uniform" else "zeros"))
        self._meta[full] = (tuple(shape[1:]), dtype)
        return full

    def pull(self, name, id_tensor):
        ids = torch.as_tensor(id_tensor).long().reshape(-1)
        own = self._owner(name, ids)
        tail, dtype = self._meta[name]
        out = torch.empty((ids.shape[0],) + tuple(tail), dtype=dtype)
        reqs, sel = [], []
        for p in torch.unique(own).tolist():
            m = torch.nonzero(own == p).squeeze(1)
            reqs.append((p, self._dg.SVC_PULL, (name, ids[m])))
            sel.append(m)
        for m, res in zip(sel, self._dg._call_many(reqs)):
            out[m] = res
        return out

    def push(self, name, id_tensor, data_tensor):
        ids = torch.as_tensor(id_tensor).long().reshape(-1)
        own = self._owner(name, ids)
        reqs = []
        for p in torch.unique(own).tolist():
            m = torch.nonzero(own == p).squeeze(1)
            reqs.append((p, self._dg.SVC_PUSH, (name, ids[m], data_tensor[m])))
        self._dg._call_many(reqs)

    def barrier(self):
        self._dg._call_many([(sid, self._dg.SVC_BARRIER, None) for sid in range(self.num_servers)])

    # ---- reference kvstore.py KVClient extras ------------------------------------------------------------------
    @property
    def client_id(self):
        c = self._dg._CTX.client
        return 0 if c is None else c.client_id

    @property
    def machine_id(self):
        return 0 if self._gpb is None else self._gpb.partid

    @property
    def role(self):
        return "default"

    @property
    def all_possible_part_policy(self):
        doms = sorted({n.rsplit("~", 1)[0] for n in self._meta})
        return {d: PartitionPolicy(d, self._gpb) for d in doms}

    def gdata_name_list(self):
        return self.data_name_list()

    def get_partid(self, name, id_tensor):
        """Owning partition of every id of tensor ``name``."""
        return self._owner(name, torch.as_tensor(id_tensor).long().reshape(-1))

    def count_nonzero(self, name):
        tail, dtype = self._meta[name]
        dom = name.split("~", 1)[0]
        n = self._gpb._num_nodes() if dom == "node" and hasattr(self._gpb, "_num_nodes") else None
        if n is None:
            m = self._gpb._node_map if dom == "node" else self._gpb._edge_map
            n = int(sum(int((v[:, 1] - v[:, 0]).sum()) for v in m.values()))
        tot = 0
        for lo in range(0, n, 1 << 20):
            tot += int((self.pull(name, torch.arange(lo, min(n, lo + (1 << 20)))) != 0).sum())
        return tot

    def delete_data(self, name):
        """Forget tensor ``name`` on this client (servers keep non-persistent data until shutdown)."""
        self._meta.pop(name, None)

    def union(self, operand1_name, operand2_name, output_name):
        """output = operand1 | operand2 (element-wise OR of two boolean / integer masks stored in the KVStore)."""
        tail, dtype = self._meta[operand1_name]
        dom = operand1_name.split("~", 1)[0]
        m = self._gpb._node_map if dom == "node" else self._gpb._edge_map
        n = int(sum(int((v[:, 1] - v[:, 0]).sum()) for v in m.values()))
        if output_name not in self._meta:
            self.init_data(output_name, (n,) + tuple(tail), dtype,
                           PartitionPolicy(operand1_name.rsplit("~", 1)[0], self._gpb))
        for lo in range(0, n, 1 << 20):
            ids = torch.arange(lo, min(n, lo + (1 << 20)))
            self.push(output_name if "~" in output_name else [k for k in self._meta if k.endswith("~" + output_name)][0], ids,
                      self.pull(operand1_name, ids) | self.pull(operand2_name, ids))

    def register_pull_handler(self, name, func):
        if self._dg._CTX.mode != "standalone":
            raise DGLError("pull handlers must be registered on the servers")
        for st in self._dg._CTX.local_servers:
            getattr(st, "pull_handlers", st.handlers)[("pull", name)] = func

    def register_push_handler(self, name, func):
        if self._dg._CTX.mode != "standalone":
            raise DGLError("push handlers must be registered on the servers (KVServer.register_push_handler)")
        for st in self._dg._CTX.local_servers:
            st.handlers[name] = func


# ---- process-wide client handle + default handlers ----------------------------------------------------------------------
_CLIENT = {"kv": None}


def default_push_handler(target, name, id_tensor, data_tensor):
    """Plain assignment of pushed rows (reference kvstore.py:39-60): ``target[name][ids] = data``."""
    target[name][id_tensor.long()] = data_tensor.to(target[name].dtype)


def default_pull_handler(target, name, id_tensor):
    """Row lookup answering a pull (reference kvstore.py:63-80)."""
    return target[name][id_tensor.long()]


def init_kvstore(ip_config=None, num_servers=1, role="default"):
    """Create the process-wide KVClient once (standalone and distributed mode share the class here)."""
    if _CLIENT["kv"] is None:
        _CLIENT["kv"] = KVClient(ip_config, num_servers, role)
    return _CLIENT["kv"]


def close_kvstore():
    _CLIENT["kv"] = None


def get_kvstore():
    return _CLIENT["kv"]


def __getattr__(name):
    # class-based request / response types of this service live in messages.py (lazy: avoids an import cycle)
    if name.startswith("__"):
        raise AttributeError(name)
    from . import messages
    if name in messages.__all__:
        return getattr(messages, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
