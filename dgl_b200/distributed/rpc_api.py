"""Class-based RPC API of the reference on top of the native transport (parity:
/root/reference/python/dgl/distributed/rpc.py: Request/Response :396-560, register_service :366, send_request :641,
recv_response :790, remote_call :873, RPCMessage :562, serialize_to_payload :455, context getters/setters :140-360).

A *service* is a pair of classes: the request carries its arguments and implements ``process_request(server_state)``,
which runs on the server and returns a response object (or ``None`` for one-way requests). Requests travel through
``rpc.Endpoint`` (csrc/cpu/rpc.cc); tensors are split out of the pickle stream by :func:`serialize_to_payload` so they
are written to the socket from their own storage instead of being copied into the pickle buffer."""
from __future__ import annotations

import io
import pickle
from collections import deque

import torch

from ..base import DGLError
from . import rpc as _rpc

__all__ = ["Request", "Response", "register_service", "get_service_property", "RPCMessage", "serialize_to_payload",
           "deserialize_from_payload", "send_request", "send_request_to_machine", "send_requests_to_machine",
           "recv_response", "recv_responses", "remote_call", "remote_call_to_machine", "send_rpc_message",
           "recv_rpc_message", "send_response", "recv_request", "client_barrier", "register_client", "get_client",
           "reset", "set_rank", "get_rank", "set_machine_id", "get_machine_id", "set_num_machines", "get_num_machines",
           "set_num_server", "get_num_server", "set_num_client", "get_num_client", "set_num_server_per_machine",
           "get_num_server_per_machine", "set_group_id", "get_group_id", "set_msg_seq", "get_msg_seq", "incr_msg_seq",
           "create_sender", "create_receiver", "finalize_sender", "finalize_receiver", "connect_receiver",
           "connect_receiver_finalize", "wait_for_senders", "finalize_server", "fast_pull",
           "copy_data_to_shared_memory", "register_sig_handler", "PlaceHolder", "ShutDownRequest",
           "ClientRegisterRequest", "ClientRegisterResponse", "GetNumberClientsRequest", "GetNumberClientsResponse",
           "ClientBarrierRequest", "ClientBarrierResponse", "CLIENT_REGISTER", "SHUT_DOWN_SERVER", "GET_NUM_CLIENT",
           "CLIENT_BARRIER"]

# ---- service registry ------------------------------------------------------------------------------------------
_REQ2ID, _ID2PAIR = {}, {}


class Request:
    """Base of request classes. Subclasses hold their arguments as attributes (pickled whole; define
    ``__getstate__``/``__setstate__`` to control it, as in the reference) and implement :meth:`process_request`."""

    def process_request(self, server_state):
        raise NotImplementedError

    @property
    def service_id(self):
        sid = _REQ2ID.get(type(self))
        if sid is None:
            raise DGLError(f"Request class {type(self).__name__} has not been registered as a service.")
        return sid


class Response:
    """Base of response classes."""

    @property
    def service_id(self):
        for sid, (_, res) in _ID2PAIR.items():
            if res is type(self):
                return sid
        raise DGLError(f"Response class {type(self).__name__} has not been registered as a service.")


def register_service(service_id, req_cls, res_cls=None):
    """Register ``service_id`` -> (request class, response class). A plain callable ``handler(state, payload)`` in
    second position keeps the functional form of :func:`rpc.register_service`."""
    if not (isinstance(req_cls, type) and issubclass(req_cls, Request)):
        return _rpc.register_service(service_id, req_cls)
    _REQ2ID[req_cls] = service_id
    _ID2PAIR[service_id] = (req_cls, res_cls)
    _rpc._SERVICES[service_id] = lambda state, req: req.process_request(state)


def get_service_property(service_id):
    """(request class, response class) of a service id, or ``None``."""
    return _ID2PAIR.get(service_id)


# ---- process-level context ---------------------------------------------------------------------------------------
_CTX = {"rank": -1, "machine_id": -1, "num_machines": 0, "num_server": 0, "num_client": 0, "num_server_per_machine": 1,
        "group_id": 0, "msg_seq": 0}
_PENDING = deque()        # seqs of requests in flight, in send order


def reset():
    """Forget rank/topology/sequence state (called at process start-up and shutdown)."""
    _CTX.update(rank=-1, machine_id=-1, num_machines=0, num_server=0, num_client=0, num_server_per_machine=1,
                group_id=0, msg_seq=0)
    _PENDING.clear()


def _accessor(key):
    def setter(value):
        _CTX[key] = int(value)

    def getter():
        return _CTX[key]
    return setter, getter


set_rank, get_rank = _accessor("rank")
set_machine_id, get_machine_id = _accessor("machine_id")
set_num_machines, get_num_machines = _accessor("num_machines")
set_num_server, get_num_server = _accessor("num_server")
set_num_client, get_num_client = _accessor("num_client")
set_num_server_per_machine, get_num_server_per_machine = _accessor("num_server_per_machine")
set_group_id, get_group_id = _accessor("group_id")
set_msg_seq, get_msg_seq = _accessor("msg_seq")


def incr_msg_seq():
    """Next message sequence number."""
    _CTX["msg_seq"] += 1
    return _CTX["msg_seq"]


# ---- payload (de)serialisation -------------------------------------------------------------------------------------
class _TensorSplitter(pickle.Pickler):
    def __init__(self, file, tensors):
        super().__init__(file, protocol=pickle.HIGHEST_PROTOCOL)
        self._tensors = tensors

    def persistent_id(self, obj):
        if torch.is_tensor(obj):
            self._tensors.append(obj.detach().cpu().contiguous())
            return len(self._tensors) - 1
        return None


class _TensorJoiner(pickle.Unpickler):
    def __init__(self, file, tensors):
        super().__init__(file)
        self._tensors = tensors

    def persistent_load(self, pid):
        return self._tensors[pid]


def serialize_to_payload(serializable):
    """(bytes, [tensors]): the object's pickle with every tensor replaced by its position in the tensor list."""
    tensors, buf = [], io.BytesIO()
    _TensorSplitter(buf, tensors).dump(serializable)
    return buf.getvalue(), tensors


def deserialize_from_payload(cls, data, tensors):
    """Inverse of :func:`serialize_to_payload`; ``cls`` is checked against the decoded object's type."""
    obj = _TensorJoiner(io.BytesIO(data), tensors).load()
    if cls is not None and not isinstance(obj, cls):
        raise DGLError(f"payload decodes to {type(obj).__name__}, expected {cls.__name__}")
    return obj


class RPCMessage:
    """Wire-level message: addressing (service, sequence, client, server, group) + payload bytes + tensors."""

    def __init__(self, service_id, msg_seq, client_id, server_id, data, tensors, group_id=0):
        self.service_id, self.msg_seq, self.client_id, self.server_id = service_id, msg_seq, client_id, server_id
        self.data, self.tensors, self.group_id = data, list(tensors), group_id


# ---- client side ---------------------------------------------------------------------------------------------------
def _client():
    c = _rpc.get_default_client()
    if c is None:
        raise DGLError("No RPC client in this process; call connect_to_server() / dgl.distributed.initialize() first.")
    return c


def get_client():
    return _rpc.get_default_client()


def register_client(client):
    """Install ``client`` (an :class:`rpc.RPCClient`) as this process's default."""
    _rpc._DEFAULT_CLIENT[0] = client


def _expects_response(request):
    pair = _ID2PAIR.get(request.service_id)
    return pair is not None and pair[1] is not None


def send_request(target, request):
    """Send ``request`` to server ``target``; its response (if the service has one) is collected by
    :func:`recv_response` in send order."""
    oneway = not _expects_response(request)
    seq = _client().send_request(int(target), request.service_id, request, oneway=oneway)
    incr_msg_seq()
    if not oneway:
        _PENDING.append(seq)


def _server_of_machine(machine_id):
    per = max(get_num_server_per_machine(), 1)
    return int(machine_id) * per + (max(get_rank(), 0) % per)


def send_request_to_machine(target, request):
    """Send to one of the servers of machine ``target`` (spread by this client's rank)."""
    send_request(_server_of_machine(target), request)


def send_requests_to_machine(target_and_requests):
    """Several (machine id, request) pairs; returns the number of responses to expect."""
    n = 0
    for target, request in target_and_requests:
        send_request_to_machine(target, request)
        n += int(_expects_response(request))
    return n


def recv_response(timeout=0):
    """Next outstanding response (send order). ``timeout`` in ms, 0 = wait indefinitely."""
    if not _PENDING:
        raise DGLError("recv_response() called with no request in flight")
    seq = _PENDING.popleft()
    return _client().recv_response(seq, timeout_ms=timeout if timeout > 0 else 3600_000)


def recv_responses(msgseq2pos, timeout=0):
    """Receive ``len(msgseq2pos)`` responses and place them at the given positions."""
    out = [None] * (max(msgseq2pos.values()) + 1 if msgseq2pos else 0)
    for _, pos in sorted(msgseq2pos.items()):
        out[pos] = recv_response(timeout)
    return out


def remote_call(target_and_requests, timeout=0):
    """Send every (server id, request) pair, then gather the responses in the same order (``None`` for one-way)."""
    c = _client()
    seqs = []
    for target, request in target_and_requests:
        oneway = not _expects_response(request)
        seq = c.send_request(int(target), request.service_id, request, oneway=oneway)
        incr_msg_seq()
        seqs.append(None if oneway else seq)
    wait = timeout if timeout > 0 else 3600_000
    return [None if s is None else c.recv_response(s, timeout_ms=wait) for s in seqs]


def remote_call_to_machine(target_and_requests, timeout=0):
    return remote_call([(_server_of_machine(t), r) for t, r in target_and_requests], timeout)


def send_rpc_message(msg, target):
    """Low-level send of an :class:`RPCMessage` to server ``target`` (payload already serialised)."""
    req = deserialize_from_payload(None, msg.data, msg.tensors)
    seq = _client().send_request(int(target), msg.service_id, req, oneway=False)
    _PENDING.append(seq)


def recv_rpc_message(timeout=0):
    """Low-level receive: the next response wrapped as an :class:`RPCMessage`."""
    res = recv_response(timeout)
    data, tensors = serialize_to_payload(res)
    return RPCMessage(getattr(res, "service_id", -1) if isinstance(res, Response) else -1, get_msg_seq(),
                      getattr(_client(), "client_id", -1), -1, data, tensors, get_group_id())


def client_barrier():
    """Block until every client of this group reached the barrier (coordinated by server 0)."""
    send_request(0, ClientBarrierRequest(get_group_id()))
    res = recv_response()
    if getattr(res, "msg", "") != "barrier":
        raise DGLError("client barrier failed")


# ---- server side ---------------------------------------------------------------------------------------------------
def recv_request(timeout=0):
    """The server loop (``rpc.RPCServer.serve_forever``) owns the receive side: it decodes each request and calls
    ``process_request``. Stand-alone polling is therefore not part of this design."""
    raise DGLError("recv_request() is handled inside RPCServer.serve_forever(); register a service instead.")


def send_response(target, response, group_id=0):
    raise DGLError("send_response() is handled inside RPCServer.serve_forever(): return the response object from "
                   "Request.process_request().")


def finalize_server():
    """Nothing to flush: ``RPCServer.serve_forever`` closes its endpoint when every client said goodbye."""
    reset()


# ---- transport lifecycle (sender/receiver pairs of the reference collapse into one Endpoint per process) ------------
def create_sender(max_queue_size=None, net_type="socket"):
    return None


def create_receiver(max_queue_size=None, net_type="socket"):
    return None


def finalize_sender():
    c = _rpc.get_default_client()
    if c is not None:
        c.shutdown()
        _rpc._DEFAULT_CLIENT[0] = None


def finalize_receiver():
    return None


def connect_receiver(ip_addr, port, recv_id, group_id=-1):
    """Connections are made by :func:`rpc.connect_to_server` for all servers at once."""
    return True


def connect_receiver_finalize(max_try_times=None):
    return True


def wait_for_senders(ip_addr, port, num_senders, blocking=True):
    return None


def fast_pull(name, id_tensor, part_id, service_id, machine_count, group_count, machine_id, client_id, local_data,
              policy):
    """Gather rows ``id_tensor`` of the distributed tensor ``name``: local rows straight from ``local_data``, remote
    rows with one pull request per owning machine (reference rpc.py:1161 does this in C++)."""
    ids = torch.as_tensor(id_tensor)
    owner = torch.as_tensor(part_id)
    out = torch.empty((ids.shape[0],) + tuple(local_data.shape[1:]), dtype=local_data.dtype)
    mine = owner == machine_id
    if bool(mine.any()):
        out[mine] = local_data[policy.to_local(ids[mine]) if hasattr(policy, "to_local") else ids[mine]]
    pair = get_service_property(service_id)
    if pair is None:
        if bool((~mine).any()):
            raise DGLError(f"service {service_id} is not registered")
        return out
    reqs, sel = [], []
    for m in torch.unique(owner[~mine]).tolist():
        pick = owner == m
        reqs.append((m, pair[0](name, ids[pick])))
        sel.append(pick)
    for pick, res in zip(sel, remote_call_to_machine(reqs)):
        out[pick] = res.data_tensor if hasattr(res, "data_tensor") else res
    return out


def copy_data_to_shared_memory(dst, source):
    """Copy into a shared-memory tensor (row by row semantics of ``copy_``)."""
    dst.copy_(source)


def register_sig_handler():
    """SIGINT/SIGTERM close the client's connections before the default handler runs."""
    import signal

    def handler(signum, frame):
        finalize_sender()
        signal.signal(signum, signal.SIG_DFL)
        signal.raise_signal(signum)

    for sig in (signal.SIGINT, signal.SIGTERM):
        try:
            signal.signal(sig, handler)
        except ValueError:      # not the main thread
            pass


class PlaceHolder:
    """Stands in for a tensor inside a pickled message (the reference swaps tensors for these when it splits the
    payload); :func:`serialize_to_payload` uses pickle persistent ids for the same purpose."""

    def __init__(self, idx=0):
        self.idx = idx


# ---- built-in services -----------------------------------------------------------------------------------------------
CLIENT_REGISTER, SHUT_DOWN_SERVER, GET_NUM_CLIENT, CLIENT_BARRIER = 22451, 22452, 22453, 22454


def _state_dict(state):
    """Per-server scratch dict living on the server-state object (or the dict itself)."""
    if isinstance(state, dict):
        return state.setdefault("_rpc_builtin", {})
    if not hasattr(state, "_rpc_builtin"):
        try:
            state._rpc_builtin = {}
        except AttributeError:
            return _FALLBACK_STATE
    return state._rpc_builtin


_FALLBACK_STATE = {}


class ClientRegisterResponse(Response):
    def __init__(self, client_id):
        self.client_id = client_id


class ClientRegisterRequest(Request):
    """Announce a client (ip address string); the server answers with a dense client id."""

    def __init__(self, ip_addr):
        self.ip_addr = ip_addr

    def process_request(self, server_state):
        st = _state_dict(server_state)
        ids = st.setdefault("clients", [])
        ids.append(self.ip_addr)
        return ClientRegisterResponse(len(ids) - 1)


class ShutDownRequest(Request):
    """Ask the server to stop serving (one-way)."""

    def __init__(self, client_id, force_shutdown_server=False):
        self.client_id, self.force_shutdown_server = client_id, force_shutdown_server

    def process_request(self, server_state):
        _state_dict(server_state)["shutdown"] = True
        return None


class GetNumberClientsResponse(Response):
    def __init__(self, num_client):
        self.num_client = num_client


class GetNumberClientsRequest(Request):
    def __init__(self, client_id=0):
        self.client_id = client_id

    def process_request(self, server_state):
        st = _state_dict(server_state)
        return GetNumberClientsResponse(max(len(st.get("clients", [])), get_num_client()))


class ClientBarrierResponse(Response):
    def __init__(self, msg="barrier"):
        self.msg = msg


class ClientBarrierRequest(Request):
    """Single-server barrier token. With one client per group it completes immediately; multi-client groups go through
    ``KVClient.barrier`` (kvstore.py), which counts arrivals on server 0."""

    def __init__(self, group_id=0, msg="barrier"):
        self.group_id, self.msg = group_id, msg

    def process_request(self, server_state):
        st = _state_dict(server_state)
        st["barrier_count"] = st.get("barrier_count", 0) + 1
        return ClientBarrierResponse(self.msg)


register_service(CLIENT_REGISTER, ClientRegisterRequest, ClientRegisterResponse)
register_service(SHUT_DOWN_SERVER, ShutDownRequest, None)
register_service(GET_NUM_CLIENT, GetNumberClientsRequest, GetNumberClientsResponse)
register_service(CLIENT_BARRIER, ClientBarrierRequest, ClientBarrierResponse)
