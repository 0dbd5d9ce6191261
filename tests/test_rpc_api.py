"""Class-based RPC API (reference python/dgl/distributed/rpc.py; tests/python/distributed/test_rpc.py) over the native
transport: two server threads, one client."""
import socket
import threading

import pytest
import torch

from dgl_b200.distributed import rpc
from dgl_b200.distributed import rpc_api as A


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class HelloResponse(A.Response):
    def __init__(self, text, t):
        self.text, self.t = text, t


class HelloRequest(A.Request):
    def __init__(self, name, t):
        self.name, self.t = name, t

    def process_request(self, state):
        state["calls"] = state.get("calls", 0) + 1
        return HelloResponse("hi " + self.name, self.t * 2)


class Fire(A.Request):
    def __init__(self, v):
        self.v = v

    def process_request(self, state):
        state["fired"] = self.v


class Boom(A.Request):
    def process_request(self, state):
        raise ValueError("server side failure")


class BoomRes(A.Response):
    pass


@pytest.mark.timeout(120)
def test_request_response_services_end_to_end():
    A.register_service(901, HelloRequest, HelloResponse)
    A.register_service(902, Fire, None)
    assert A.get_service_property(901) == (HelloRequest, HelloResponse) and HelloRequest("x", 0).service_id == 901
    assert HelloResponse("x", 0).service_id == 901 and A.get_service_property(12345) is None
    ports = [_free_port(), _free_port()]
    cfg = [("127.0.0.1", p) for p in ports]
    states = [{}, {}]
    threads = [threading.Thread(target=rpc.start_server, args=(i, cfg, 2, 1, states[i]), daemon=True) for i in range(2)]
    for t in threads:
        t.start()
    rpc.connect_to_server(cfg, 2)
    try:
        assert A.get_client() is rpc.get_default_client()
        A.send_request(0, HelloRequest("a", torch.ones(2)))
        A.send_request(1, HelloRequest("b", torch.ones(1)))
        r0, r1 = A.recv_response(), A.recv_response()
        assert r0.text == "hi a" and r0.t.tolist() == [2.0, 2.0] and r1.text == "hi b"
        out = A.remote_call([(1, HelloRequest("c", torch.zeros(1))), (0, Fire(7)), (0, HelloRequest("d", torch.zeros(1)))])
        assert [getattr(o, "text", None) for o in out] == ["hi c", None, "hi d"]
        assert states[0]["fired"] == 7 and states[0]["calls"] == 2 and states[1]["calls"] == 2
        A.set_num_server_per_machine(1)
        A.set_rank(0)
        assert A.remote_call_to_machine([(1, HelloRequest("m", torch.zeros(1)))])[0].text == "hi m"
        assert A.send_requests_to_machine([(0, HelloRequest("n", torch.zeros(1))), (1, Fire(1))]) == 1
        assert A.recv_response().text == "hi n"
        assert A.remote_call([(0, A.ClientRegisterRequest("1.2.3.4"))])[0].client_id == 0
        assert A.remote_call([(0, A.GetNumberClientsRequest())])[0].num_client == 1
        A.client_barrier()
        with pytest.raises(Exception):
            A.recv_response()                      # nothing in flight

        class Unregistered(A.Request):
            pass

        with pytest.raises(Exception):
            A.send_request(0, Unregistered())
        # server-side exceptions travel back to the caller
        A.register_service(903, Boom, BoomRes)
        with pytest.raises(ValueError, match="server side failure"):
            A.remote_call([(0, Boom())])
    finally:
        A.finalize_sender()
    for t in threads:
        t.join(20)
    assert not any(t.is_alive() for t in threads)
    assert rpc.get_default_client() is None


def test_payload_split_and_context_accessors():
    data, tensors = A.serialize_to_payload({"x": torch.arange(3), "y": [1, torch.ones(2)]})
    assert len(tensors) == 2 and isinstance(data, bytes) and len(data) < 200      # tensors are not inside the pickle
    back = A.deserialize_from_payload(dict, data, tensors)
    assert back["x"].tolist() == [0, 1, 2] and back["y"][1].tolist() == [1.0, 1.0]
    with pytest.raises(Exception):
        A.deserialize_from_payload(list, data, tensors)
    msg = A.RPCMessage(5, 1, 2, 3, data, tensors, group_id=4)
    assert (msg.service_id, msg.msg_seq, msg.client_id, msg.server_id, msg.group_id) == (5, 1, 2, 3, 4)
    A.reset()
    assert A.get_rank() == -1 and A.get_msg_seq() == 0
    for setter, getter, v in [(A.set_rank, A.get_rank, 3), (A.set_machine_id, A.get_machine_id, 1),
                              (A.set_num_machines, A.get_num_machines, 2), (A.set_num_server, A.get_num_server, 4),
                              (A.set_num_client, A.get_num_client, 8), (A.set_group_id, A.get_group_id, 5),
                              (A.set_num_server_per_machine, A.get_num_server_per_machine, 2)]:
        setter(v)
        assert getter() == v
    assert A.incr_msg_seq() == 1 and A.incr_msg_seq() == 2
    A.set_msg_seq(10)
    assert A.get_msg_seq() == 10
    # names are reachable from the rpc module too, as in the reference's single file
    assert rpc.Request is A.Request and rpc.remote_call is A.remote_call and rpc.get_num_server() == 4
    A.reset()
    # local rows of fast_pull need no network
    local = torch.arange(12.0).view(6, 2)
    got = A.fast_pull("w", torch.tensor([4, 1]), torch.tensor([0, 0]), 999, 1, 1, 0, 0, local, None)
    assert got.tolist() == [[8.0, 9.0], [2.0, 3.0]]
    shared = torch.zeros(2, 2)
    A.copy_data_to_shared_memory(shared, torch.ones(2, 2))
    assert shared.sum() == 4
