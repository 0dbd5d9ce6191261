"""Reference-named helpers around the distributed core (reference python/dgl/distributed/{id_map,graph_partition_book,
kvstore,rpc_client,graph_services,partition,dist_context}.py; tests/python/distributed/test_partition.py::test_hetero_*)."""
import os
import tempfile

import numpy as np
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200 import distributed as dd


def test_id_map_decodes_homogeneous_ids():
    m = dd.IdMap({"T0": np.array([[0, 100], [200, 300]]), "T1": np.array([[100, 200], [300, 400]])})
    t, p = m(torch.tensor([90, 201, 150, 399, 0]))
    assert t.tolist() == [0, 0, 1, 1, 0] and p.tolist() == [90, 101, 50, 199, 0]
    # exhaustive: per-type ids enumerate 0..199 for each type
    t, p = m(torch.arange(400))
    for k in (0, 1):
        assert sorted(p[t == k].tolist()) == list(range(200))
    with pytest.raises(dgl.DGLError):
        m(torch.tensor([400]))
    e, ids = m(torch.tensor([], dtype=torch.long))
    assert e.numel() == 0
    t0, p0 = dd.IdMap({})(torch.tensor([3, 4]))
    assert t0.tolist() == [0, 0] and p0.tolist() == [3, 4]
    # uneven partitions, three types given out of layout order
    r = {"b": np.array([[5, 9], [20, 21]]), "a": np.array([[0, 5], [12, 20]]), "c": np.array([[9, 12], [21, 30]])}
    t, p = dd.IdMap(r)(torch.arange(30))
    assert t[:12].tolist() == [0] * 5 + [1] * 4 + [2] * 3 and p[12:20].tolist() == list(range(5, 13))


def test_data_names_and_policies():
    n = dd.HeteroDataName(True, "user", "feat")
    assert str(n) == "node~user~feat" and n.policy_str == "node~user" and n.is_node() and n.get_name() == "feat"
    e = dd.parse_hetero_data_name("edge~a:r:b~w")
    assert e.get_type() == ("a", "r", "b") and e.is_edge() and str(e) == "edge~a:r:b~w"
    with pytest.raises(AssertionError):
        dd.HeteroDataName(False, "r", "w")
    with pytest.raises(AssertionError):
        dd.parse_hetero_data_name("bogus")
    g = dgl.rand_graph(200, 800)
    with tempfile.TemporaryDirectory() as d:
        dd.partition_graph(g, "cg", 2, d)
        gpb = dd.load_partition_book(os.path.join(d, "cg.json"), 0)[0]
        pol = dd.NodePartitionPolicy(gpb)
        assert pol.policy_str == "node:_N" and pol.get_size() == 200 and 0 < pol.get_part_size() < 200
        assert dd.EdgePartitionPolicy(gpb).policy_str == "edge:_N:_E:_N"
        part = dd.get_node_partition_from_book(gpb, "cpu")
        assert part.num_parts() == 2 and part.array_size() == 200
        assert part.local_size(0) + part.local_size(1) == 200
        assert dd.get_shared_mem_partition_book("no_such_graph_xyz") is None
        gpb.shared_memory("cg_compat_test")
        back = dd.get_shared_mem_partition_book("cg_compat_test")
        assert back.num_partitions() == 2 and back.partid == 0
        from dgl_b200.utils.shared_mem import unlink
        for seg in ["/cg_compat_test_gpb_index"] + [f"/cg_compat_test_gpb_t{i}" for i in range(4)]:
            try:
                unlink(seg)
            except Exception:  # noqa: BLE001 - fewer segments than probed
                pass


def test_kvstore_handlers_and_misc_functions():
    store = {"w": torch.zeros(4, 2)}
    dd.default_push_handler(store, "w", torch.tensor([1, 3]), torch.ones(2, 2))
    assert dd.default_pull_handler(store, "w", torch.tensor([3, 0])).tolist() == [[1.0, 1.0], [0.0, 0.0]]
    assert dd.get_kvstore() is None
    dd.close_kvstore()
    assert "127.0.0.1" in dd.local_ip4_addr_list()
    addr = dd.get_local_usable_addr("127.0.0.1")
    assert addr.count(":") == 1 and int(addr.split(":")[1]) > 0
    assert dd.get_local_machine_id({0: [7, "127.0.0.1", 30050]}) == 7
    with pytest.raises(dgl.DGLError):
        dd.get_local_machine_id({0: [7, "203.0.113.9", 30050]})
    g = dd.merge_graphs([(torch.tensor([0]), torch.tensor([1]), torch.tensor([5])),
                         (torch.tensor([2]), torch.tensor([1]), torch.tensor([9]))], 4, exclude_edges=[9])
    assert g.num_nodes() == 4 and g.edata[dgl.EID].tolist() == [5]
    assert dd.merge_graphs([], 3).num_edges() == 0
    assert dd.is_homogeneous(["_N"], ["_E"]) and not dd.is_homogeneous(["a", "b"], ["r"])
    ip, idx, tpe, na, ea = dd.cast_various_to_minimum_dtype_gb(1, torch.tensor([0, 1, 2]), torch.tensor([1, 0]),
                                                               torch.tensor([0, 1]), ["a", "b"], ["n"], None, None)
    assert ip.dtype == torch.int32 and idx.dtype == torch.int32 and tpe.dtype == torch.uint8 and na is None and ea is None
    big = dd.cast_various_to_minimum_dtype_gb(1, torch.tensor([0, 1]), torch.tensor([0]), torch.tensor([300]),
                                              ["e"] * 400, ["n"], None, None)
    assert big[2].dtype == torch.int16
    pg = dd.process_partitions(dgl.rand_graph(10, 30), formats=["csc"])
    assert pg.num_edges() == 30
    pool, n = dd.get_sampler_pool()
    pool.set_collate_fn(lambda a, b: a + b, "dl")
    pool.submit_task("dl", (1, 2))
    assert pool.get_result("dl") == 3 and n == pool.num_workers
    pool.delete_collate_fn("dl")
    dd.finalize_worker()
    dd.join_finalize_worker()
    assert dd.MpCommand.FINALIZE_POOL == 6


def test_message_classes_run_the_server_handlers():
    """Request objects executed directly against a partition's ServerState (no network needed)."""
    from dgl_b200.distributed import dist_graph, kvstore, role
    from dgl_b200.distributed.rpc_api import get_service_property
    g = dgl.rand_graph(300, 1500)
    g.ndata["feat"] = torch.randn(300, 4)
    g.edata["w"] = torch.rand(1500)
    with tempfile.TemporaryDirectory() as d:
        dd.partition_graph(g, "mg", 2, d)
        st = dist_graph.ServerState(os.path.join(d, "mg.json"), 0)
        lo = st.nid_start
        ids = torch.arange(lo, lo + 3)
        res = kvstore.PullRequest("node~_N~feat", ids).process_request(st)
        assert res.server_id == 0 and torch.equal(res.data_tensor, st.kv["node~_N~feat"][:3])
        assert kvstore.PushRequest("node~_N~feat", ids, torch.ones(3, 4)).process_request(st) is None
        assert torch.equal(st.kv["node~_N~feat"][:3], torch.ones(3, 4))
        kvstore.RegisterPushHandlerRequest("node~_N~feat", lambda t, loc, x: t.index_add_(0, loc, x)).process_request(st)
        kvstore.PushRequest("node~_N~feat", ids, torch.ones(3, 4)).process_request(st)
        assert torch.equal(st.kv["node~_N~feat"][:3], torch.full((3, 4), 2.0))
        assert kvstore.InitDataRequest("node~_N~acc", (300, 2), torch.float32, "node:_N").process_request(st).msg == "ok"
        assert kvstore.GetPartShapeRequest("node~_N~acc").process_request(st).shape == (st.n_inner, 2)
        meta = kvstore.GetSharedDataRequest().process_request(st).meta
        assert meta["node~_N~acc"][2] == "node:_N" and "node~_N~feat" in meta
        assert kvstore.CountLocalNonzeroRequest("node~_N~acc").process_request(st).num_local_nonzero == 0
        kvstore.DeleteDataRequest("node~_N~acc").process_request(st)
        assert "node~_N~acc" not in st.kv
        assert kvstore.BarrierRequest().process_request(st).msg == "ok"
        kvstore.SendMetaToBackupRequest("x", torch.float32, (1,), "node:_N").process_request(st)
        role.RegisterRoleRequest(3, 0, "trainer").process_request(st)
        role.RegisterRoleRequest(4, 1, "trainer").process_request(st)
        assert role.GetRoleRequest().process_request(st).role == {"trainer": [(3, 0), (4, 1)]}
        assert dist_graph.InitGraphRequest("mg").process_request(st).msg == "mg"
        with pytest.raises(RuntimeError):
            dist_graph.InitGraphRequest("other").process_request(st)
        assert dist_graph.QueryIfUseGraphBoltRequest().process_request(st).use_graphbolt is False
        dist_graph.AddEdgeAttributeFromSharedMemRequest("w").process_request(st)
        assert st.g.edata["w"].shape[0] == st.g.num_edges()
        sub = dd.SamplingRequest(ids, 2).process_request(st)
        assert sub.global_dst.numel() <= 6 and torch.isin(sub.global_dst, ids).all()
        ins = dd.InSubgraphRequest(ids).process_request(st)
        assert torch.isin(ins.global_dst, ids).all()
        e = dd.EdgesRequest(torch.tensor([st.eid_start]), 5).process_request(st)
        assert e.order_id == 5 and e.global_src.numel() == 1
        # every class is registered as a service with a response class (or None for one-way pushes)
        assert get_service_property(kvstore.PullRequest("a", ids).service_id) == (kvstore.PullRequest, kvstore.PullResponse)
        assert get_service_property(kvstore.PushRequest("a", ids, ids).service_id)[1] is None
    assert dist_graph.NodeDataView is dist_graph.EdgeDataView


def test_kvstore_handles_partition_helpers_and_role_rank():
    import torch
    import dgl_b200 as dgl
    from dgl_b200.distributed import kvstore as K, partition as P, role as R
    store = {"emb": torch.zeros(4, 2)}
    K.default_push_handler(store, "emb", torch.tensor([1, 3]), torch.ones(2, 2))
    assert store["emb"].sum() == 4 and K.default_pull_handler(store, "emb", torch.tensor([3])).tolist() == [[1.0, 1.0]]
    K.close_kvstore()
    assert K.get_kvstore() is None
    assert P.is_homogeneous(["_N"], ["_E"]) and not P.is_homogeneous(["a", "b"], ["e"])
    assert R.get_rank() == R.get_trainer_rank()
    ip, ix = torch.tensor([0, 2, 5]), torch.tensor([1, 0, 1, 0, 1])
    out = P.cast_various_to_minimum_dtype_gb(2, ip, ix, torch.tensor([0, 1, 0, 1, 1]), ["a", "b"], ["n"],
                                             {dgl.NID: torch.arange(2)}, {dgl.EID: torch.arange(5), "inner_edge": torch.ones(5)},
                                             tot_edge_count=5, tot_node_count=2, edge_count=5, node_count=2)
    assert out[0].dtype == torch.int32 and out[1].dtype == torch.int32 and out[2].dtype == torch.uint8
    assert out[4]["inner_edge"].dtype == torch.uint8 and out[3][dgl.NID].dtype == torch.int32
    big = P.cast_various_to_minimum_dtype_gb(2, ip, ix, None, ["a"], ["n"], {}, {}, tot_edge_count=2 ** 33,
                                             tot_node_count=10, edge_count=5, node_count=2)
    assert big[0].dtype == torch.int64 and big[1].dtype == torch.int32 and big[2] is None
    g = dgl.rand_graph(10, 30)
    g.ndata["inner_node"] = torch.ones(10)
    assert P.process_partitions(g).ndata["inner_node"].dtype == torch.uint8
