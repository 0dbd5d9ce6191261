"""Request / response classes of the KVStore, role and DistGraph services in the reference's class-based shape (parity:
/root/reference/python/dgl/distributed/kvstore.py:30-800, role.py:20-110, dist_graph.py:60-240).

The live services of this package are the function handlers of ``dist_graph`` (``_h_pull`` ...), addressed by small
integer ids; the classes here run the *same* handlers from ``process_request`` so either calling convention reaches
the same server-side code. They are registered under the reference's wire ids (service_ids.py) with :func:`rpc_api.register_service`."""
from __future__ import annotations

import torch

from . import dist_graph as dg
from .rpc_api import Request, Response, register_service

__all__ = []


def _export(cls):
    __all__.append(cls.__name__)
    return cls


def _ok_response(name, doc):
    cls = type(name, (Response,), {"__init__": lambda self, msg="ok": setattr(self, "msg", msg), "__doc__": doc})
    return _export(cls)


# ---- KVStore ------------------------------------------------------------------------------------------------------------
@_export
class PullResponse(Response):
    """Rows answered by a :class:`PullRequest`."""

    def __init__(self, server_id, data_tensor):
        self.server_id, self.data_tensor = server_id, data_tensor


@_export
class PullRequest(Request):
    """Read rows ``id_tensor`` (global ids owned by the target server) of tensor ``name``."""

    def __init__(self, name, id_tensor):
        self.name, self.id_tensor = name, id_tensor

    def process_request(self, server_state):
        return PullResponse(server_state.part_id, dg._h_pull(server_state, (self.name, self.id_tensor)))


@_export
class PushRequest(Request):
    """Write rows (one-way; the registered push handler of ``name`` decides how to combine)."""

    def __init__(self, name, id_tensor, data_tensor):
        self.name, self.id_tensor, self.data_tensor = name, id_tensor, data_tensor

    def process_request(self, server_state):
        dg._h_push(server_state, (self.name, self.id_tensor, self.data_tensor))


InitDataResponse = _ok_response("InitDataResponse", "Acknowledges :class:`InitDataRequest`.")


@_export
class InitDataRequest(Request):
    """Create the shard of a new distributed tensor on the target server."""

    def __init__(self, name, shape, dtype, policy_str, init_func=None):
        self.name, self.shape, self.dtype, self.policy_str, self.init_func = name, tuple(shape), dtype, policy_str, init_func

    def process_request(self, server_state):
        dom, typ = (self.policy_str.split(":", 1) + ["_N"])[:2] if ":" in self.policy_str else (self.policy_str, "_N")
        init = "uniform" if self.init_func == "uniform" else "zeros"
        dg._h_init_data(server_state, (self.name, self.shape[1:], self.dtype, dom, typ, init))
        return InitDataResponse()


BarrierResponse = _ok_response("BarrierResponse", "Releases a client from :class:`BarrierRequest`.")


@_export
class BarrierRequest(Request):
    def __init__(self, role="default"):
        self.role = role

    def process_request(self, server_state):
        dg._h_barrier(server_state, self.role)
        return BarrierResponse()


RegisterPullHandlerResponse = _ok_response("RegisterPullHandlerResponse", "Acknowledges a pull-handler registration.")
RegisterPushHandlerResponse = _ok_response("RegisterPushHandlerResponse", "Acknowledges a push-handler registration.")


@_export
class RegisterPushHandlerRequest(Request):
    """Install ``push_func(target, local_ids, data)`` for tensor ``name`` on the server."""

    def __init__(self, name, push_func):
        self.name, self.push_func = name, push_func

    def process_request(self, server_state):
        server_state.handlers[self.name] = self.push_func
        return RegisterPushHandlerResponse()


@_export
class RegisterPullHandlerRequest(Request):
    """Pulls are plain row reads in this implementation; the handler is recorded for introspection."""

    def __init__(self, name, pull_func):
        self.name, self.pull_func = name, pull_func

    def process_request(self, server_state):
        server_state.handlers.setdefault("__pull__", {})[self.name] = self.pull_func
        return RegisterPullHandlerResponse()


@_export
class GetSharedDataResponse(Response):
    def __init__(self, meta):
        self.meta = meta          # {name: (shape, dtype, policy_str)}


@_export
class GetSharedDataRequest(Request):
    """Metadata of every tensor the server holds."""

    def __init__(self, msg="get_shared_data"):
        self.msg = msg

    def process_request(self, server_state):
        meta = {}
        for k, v in server_state.kv.items():
            dom, typ = server_state.policy[k]
            meta[k] = (tuple(v.shape), v.dtype, f"{dom}:{typ}")
        return GetSharedDataResponse(meta)


@_export
class GetPartShapeResponse(Response):
    def __init__(self, shape):
        self.shape = shape


@_export
class GetPartShapeRequest(Request):
    def __init__(self, name):
        self.name = name

    def process_request(self, server_state):
        return GetPartShapeResponse(tuple(server_state.kv[self.name].shape))


DeleteDataResponse = _ok_response("DeleteDataResponse", "Acknowledges :class:`DeleteDataRequest`.")


@_export
class DeleteDataRequest(Request):
    def __init__(self, name):
        self.name = name

    def process_request(self, server_state):
        server_state.kv.pop(self.name, None)
        server_state.policy.pop(self.name, None)
        return DeleteDataResponse()


@_export
class CountLocalNonzeroResponse(Response):
    def __init__(self, num_local_nonzero):
        self.num_local_nonzero = num_local_nonzero


@_export
class CountLocalNonzeroRequest(Request):
    def __init__(self, name):
        self.name = name

    def process_request(self, server_state):
        return CountLocalNonzeroResponse(int(torch.count_nonzero(server_state.kv[self.name])))


SendMetaToBackupResponse = _ok_response("SendMetaToBackupResponse", "Acknowledges :class:`SendMetaToBackupRequest`.")


@_export
class SendMetaToBackupRequest(Request):
    """Backup servers of the reference share the main server's shards through shared memory; there is one server per
    partition here, so this only records the policy."""

    def __init__(self, name, dtype, shape, policy_str, pull_handler=None, push_handler=None):
        self.name, self.dtype, self.shape, self.policy_str = name, dtype, shape, policy_str

    def process_request(self, server_state):
        if ":" in self.policy_str:
            server_state.policy.setdefault(self.name, tuple(self.policy_str.split(":", 1)))
        return SendMetaToBackupResponse()


# ---- roles --------------------------------------------------------------------------------------------------------------
RegisterRoleResponse = _ok_response("RegisterRoleResponse", "Acknowledges :class:`RegisterRoleRequest`.")


@_export
class RegisterRoleRequest(Request):
    """A client announces (client id, machine id, role name)."""

    def __init__(self, client_id, machine_id, role):
        self.client_id, self.machine_id, self.role = client_id, machine_id, role

    def process_request(self, server_state):
        roles = server_state.__dict__.setdefault("roles", {})
        roles.setdefault(self.role, set()).add((self.client_id, self.machine_id))
        return RegisterRoleResponse()


@_export
class GetRoleResponse(Response):
    def __init__(self, role):
        self.role = role          # {role name: sorted [(client id, machine id)]}


@_export
class GetRoleRequest(Request):
    def __init__(self, msg="get_role"):
        self.msg = msg

    def process_request(self, server_state):
        roles = getattr(server_state, "roles", {})
        return GetRoleResponse({k: sorted(v) for k, v in roles.items()})


# ---- DistGraph ----------------------------------------------------------------------------------------------------------
InitGraphResponse = _ok_response("InitGraphResponse", "The server holds the requested graph.")


@_export
class InitGraphRequest(Request):
    """Handshake of a ``DistGraph`` client: checks the graph name served by the target."""

    def __init__(self, graph_name):
        self._graph_name = graph_name

    def process_request(self, server_state):
        if server_state.graph_name != self._graph_name:
            raise RuntimeError(f"server holds graph {server_state.graph_name!r}, not {self._graph_name!r}")
        return InitGraphResponse(self._graph_name)


@_export
class QueryIfUseGraphBoltResponse(Response):
    def __init__(self, use_graphbolt):
        self.use_graphbolt = use_graphbolt


@_export
class QueryIfUseGraphBoltRequest(Request):
    def __init__(self, msg="query"):
        self.msg = msg

    def process_request(self, server_state):
        return QueryIfUseGraphBoltResponse(bool(getattr(server_state, "use_graphbolt", False)))


AddEdgeAttributeFromKVResponse = _ok_response("AddEdgeAttributeFromKVResponse", "Edge attribute attached.")
AddEdgeAttributeFromSharedMemResponse = _ok_response("AddEdgeAttributeFromSharedMemResponse", "Edge attribute attached.")


@_export
class AddEdgeAttributeFromKVRequest(Request):
    """Expose a KVStore edge tensor as an edge attribute of the server's sampling graph (e.g. sampling probabilities)."""

    def __init__(self, name, kv_names):
        self._name, self._kv_names = name, kv_names

    def process_request(self, server_state):
        key = self._kv_names[0] if isinstance(self._kv_names, (list, tuple)) else self._kv_names
        data = server_state.kv[key]
        n = server_state.g.num_edges()
        full = torch.zeros((n,) + tuple(data.shape[1:]), dtype=data.dtype)
        full[: data.shape[0]] = data                       # inner edges are a prefix of the partition's edges
        server_state.g.edata[self._name] = full
        return AddEdgeAttributeFromKVResponse()


@_export
class AddEdgeAttributeFromSharedMemRequest(AddEdgeAttributeFromKVRequest):
    """Same effect; the reference's second variant reads the tensor from the shared-memory copy of the partition."""

    def __init__(self, name):
        super().__init__(name, "edge~_N:_E:_N~" + name)


from . import service_ids as _ids  # noqa: E402

for _sid, _req, _res in (
        (_ids.KVSTORE_PULL, PullRequest, PullResponse), (_ids.KVSTORE_PUSH, PushRequest, None),
        (_ids.INIT_DATA, InitDataRequest, InitDataResponse), (_ids.BARRIER, BarrierRequest, BarrierResponse),
        (_ids.REGISTER_PUSH, RegisterPushHandlerRequest, RegisterPushHandlerResponse),
        (_ids.REGISTER_PULL, RegisterPullHandlerRequest, RegisterPullHandlerResponse),
        (_ids.GET_SHARED, GetSharedDataRequest, GetSharedDataResponse),
        (_ids.GET_PART_SHAPE, GetPartShapeRequest, GetPartShapeResponse),
        (_ids.DELETE_DATA, DeleteDataRequest, DeleteDataResponse),
        (_ids.COUNT_LOCAL_NONZERO, CountLocalNonzeroRequest, CountLocalNonzeroResponse),
        (_ids.SEND_META_TO_BACKUP, SendMetaToBackupRequest, SendMetaToBackupResponse),
        (_ids.REGISTER_ROLE, RegisterRoleRequest, RegisterRoleResponse), (_ids.GET_ROLE, GetRoleRequest, GetRoleResponse),
        (_ids.INIT_GRAPH, InitGraphRequest, InitGraphResponse),
        (_ids.QUERY_IF_USE_GRAPHBOLT, QueryIfUseGraphBoltRequest, QueryIfUseGraphBoltResponse),
        (_ids.ADD_EDGE_ATTRIBUTE_FROM_KV, AddEdgeAttributeFromKVRequest, AddEdgeAttributeFromKVResponse),
        (_ids.ADD_EDGE_ATTRIBUTE_FROM_SHARED_MEM, AddEdgeAttributeFromSharedMemRequest,
         AddEdgeAttributeFromSharedMemResponse)):
    register_service(_sid, _req, _res)
