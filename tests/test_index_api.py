"""Index-level / legacy API adapters and small module-level helpers (reference python/dgl/{heterograph_index,
graph_index,random,init,convert,partition,core,base}.py; reference tests tests/python/common/test_heterograph-index.py,
tests/python/common/test_random.py)."""
import pickle

import networkx as nx
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import dgl_b200 as dgl
import dgl_b200.function as fn
from dgl_b200 import graph_index as GI
from dgl_b200 import heterograph_index as H


@pytest.fixture
def hg():
    return dgl.heterograph({("a", "r", "b"): ([0, 1, 1], [0, 0, 2]), ("b", "s", "a"): ([0], [1])})


def test_hetero_index_queries(hg):
    gi = hg._graph
    assert (gi.number_of_ntypes(), gi.number_of_etypes()) == (2, 2)
    assert gi.metagraph.find_edge(0) == (0, 1) and gi.metagraph.find_edge(1) == (1, 0)
    assert gi.dtype == "int64" and gi.ctx == torch.device("cpu") and gi.bits_needed(0) == 32
    assert gi.num_nodes(1) == 3 and gi.number_of_edges(0) == 3
    assert gi.successors(0, 1).tolist() == [0, 2] and gi.predecessors(0, 0).tolist() == [0, 1]
    u, v, e = gi.edges(0)
    assert (u.tolist(), v.tolist(), e.tolist()) == ([0, 1, 1], [0, 0, 2], [0, 1, 2])
    assert gi.in_degrees(0, torch.tensor([0, 1, 2])).tolist() == [2, 0, 1]
    assert gi.out_degrees(0, torch.tensor([0, 1])).tolist() == [1, 2]
    fu, fv, fe = gi.find_edges(0, torch.tensor([2]))
    assert (fu.tolist(), fv.tolist(), fe.tolist()) == ([1], [2], [2])
    assert gi.edge_ids_one(0, torch.tensor([1]), torch.tensor([2])).tolist() == [2]
    assert gi.has_edges_between(0, torch.tensor([0, 1]), torch.tensor([2, 2])).tolist() == [False, True]
    ip, idx, eid = gi.adjacency_matrix_tensors(0, False, "csr")
    assert ip.tolist() == [0, 1, 3] and idx.tolist() == [0, 0, 2] and eid.tolist() == [0, 1, 2]
    row, col = gi.adjacency_matrix_tensors(0, True, "coo")
    assert row.tolist() == [0, 0, 2] and col.tolist() == [0, 1, 1]
    assert gi.adjacency_matrix_scipy(0, False, "csr").shape == (3, 2)
    assert gi.adjacency_matrix_scipy(0, True, "coo", return_edge_ids=True).data.tolist() == [0, 1, 2]
    assert gi.get_csr_shuffle_order(0).tolist() == [0, 1, 2]
    assert not gi.is_multigraph() and gi.is_readonly() and not gi.is_metagraph_unibipartite()
    assert gi.asbits(32).dtype == "int32" and gi.copy_to("cpu").ctx.type == "cpu"
    assert gi.get_relation_graph(1).num_edges(0) == 1
    assert sorted(gi.reverse().graph.canonical_etypes) == [("a", "s", "b"), ("b", "r", "a")]


def test_hetero_index_subgraphs_and_creators(hg):
    gi = hg._graph
    sg = gi.node_subgraph([torch.tensor([1]), torch.tensor([0, 2])])
    assert sg.graph.num_edges(0) == 2 and sg.induced_edges[0].tolist() == [1, 2] and sg.induced_edges[1].tolist() == [0]
    assert [n.tolist() for n in sg.induced_nodes] == [[1], [0, 2]]
    es = gi.edge_subgraph([torch.tensor([0]), torch.tensor([], dtype=torch.int64)], False)
    assert es.graph.num_nodes(0) == 1 and es.induced_edges[0].tolist() == [0]
    keep = gi.edge_subgraph([torch.tensor([0]), torch.tensor([], dtype=torch.int64)], True)
    assert keep.graph.num_nodes(0) == 2 and keep.graph.num_nodes(1) == 3
    mg, nt, et, ce = H.create_metagraph_index(["a", "b"], [("a", "r", "b")])
    assert mg.num_nodes() == 2 and mg.num_edges() == 1 and et == ["r"]
    ug = H.create_unitgraph_from_coo(2, 2, 3, torch.tensor([0, 1]), torch.tensor([2, 0]))
    hi = H.create_heterograph_from_relations(mg, [ug], None, nt, ce)
    assert hi.graph.num_nodes("b") == 3 and hi.graph.num_edges() == 2
    ug2 = H.create_unitgraph_from_csr(1, 3, 3, torch.tensor([0, 1, 2, 2]), torch.tensor([1, 2]), torch.tensor([], dtype=torch.int64))
    assert ug2.edges()[0].tolist() == [0, 1] and ug2.edges()[1].tolist() == [1, 2]
    only = H.create_unitgraph_from_coo(1, 3, 3, torch.tensor([0]), torch.tensor([1]), formats=["csr"])
    assert only.allowed_formats() == ["csr"] or set(only.allowed_formats()) == {"csr"}
    b = H.disjoint_union(None, [dgl.graph(([0], [1]))._graph, dgl.graph(([0, 1], [1, 2]))._graph])
    assert b.graph.batch_size == 2
    parts = H.disjoint_partition(b, [[2, 3]], [[1, 2]])
    assert [p.graph.num_edges() for p in parts] == [1, 2]
    sl = H.slice_gidx(b, [3], [2], [2], [1])
    assert sl.graph.num_nodes() == 3 and sl.graph.num_edges() == 2


def test_legacy_graph_index_mutation_and_queries():
    gi = GI.create_graph_index(None, False)
    gi.add_nodes(4)
    gi.add_edge(0, 1)
    gi.add_edges([1, 2], [2, 3])
    gi.add_edges(3, [0, 1])
    assert (gi.num_nodes(), gi.num_edges()) == (4, 5)
    assert gi.has_node(3) and not gi.has_node(4) and gi.has_nodes([0, 9]).tolist() == [1, 0]
    assert gi.has_edge_between(0, 1) and gi.has_edges_between([0, 1], [2, 2]).tolist() == [0, 1]
    assert gi.edge_id(3, 0).tolist() == [3] and gi.find_edge(2) == (2, 3)
    assert gi.in_degree(1) == 2 and gi.out_degrees([3]).tolist() == [2]
    assert gi.predecessors(1).tolist() == [0, 3] and gi.successors(3).tolist() == [0, 1]
    u, v, e = gi.edges("srcdst")
    assert u.tolist() == [0, 1, 2, 3, 3] and v.tolist() == [1, 2, 3, 0, 1]
    s = gi.node_subgraph([1, 2, 3])
    assert s.graph.num_edges() == 3 and s.induced_edges.tolist() == [1, 2, 4] and s.parent is gi
    assert gi.node_halo_subgraph([1], 1).induced_nodes.tolist() == [1, 0, 3]
    es = gi.edge_subgraph([0, 4])
    assert es.induced_nodes.tolist() == [0, 1, 3]
    assert gi.edge_subgraph([0], preserve_nodes=True).graph.num_nodes() == 4
    assert pickle.loads(pickle.dumps(gi)).num_edges() == 5
    assert gi.to_networkx().number_of_edges() == 5 and gi.line_graph().num_nodes() == 5
    assert gi.nbits() == 64 and gi.asbits(32).dtype == "int32" and gi.bits_needed() == 32
    assert gi.adjacency_matrix_scipy(False, "csr").nnz == 5 and gi.get_csr_shuffle_order().numel() == 5
    ro = gi.to_immutable()
    assert ro.is_readonly()
    with pytest.raises(dgl.DGLError):
        ro.add_nodes(1)
    with pytest.raises(dgl.DGLError):
        GI.create_graph_index(None, True)
    gi.clear()
    assert gi.num_nodes() == 0


def test_legacy_graph_index_creators():
    assert GI.from_edge_list([(0, 1), (1, 2)], True).num_nodes() == 3
    assert GI.from_edge_list(([0, 1], [1, 2]), True).num_edges() == 2
    assert GI.create_graph_index(sp.eye(3).tocoo(), True).num_edges() == 3
    assert GI.create_graph_index(nx.path_graph(3), True).num_edges() == 4
    assert GI.from_coo(5, [0], [4], True).num_nodes() == 5
    out = GI.from_csr([0, 1, 2, 2], [1, 2], "out").edges()
    assert out[0].tolist() == [0, 1] and out[1].tolist() == [1, 2]
    inn = GI.from_csr([0, 1, 2, 2], [1, 2], "in").edges()
    assert inn[0].tolist() == [1, 2] and inn[1].tolist() == [0, 1]
    assert GI.map_to_subgraph_nid([5, 2, 9], [9, 5]).tolist() == [2, 0]
    with pytest.raises(dgl.DGLError):
        GI.map_to_subgraph_nid([5, 2, 9], [3])
    assert GI.transform_ids(torch.tensor([7, 8, 9]), [2, 0]).tolist() == [9, 7]
    u = GI.disjoint_union([GI.from_edge_list([(0, 1)], True), GI.from_edge_list([(0, 1), (1, 2)], True)])
    assert u.num_nodes() == 5 and [p.num_edges() for p in GI.disjoint_partition(u, [2, 3])] == [1, 2]
    same = GI.disjoint_union([GI.from_edge_list([(0, 1)], True)] * 3)
    assert [p.num_nodes() for p in GI.disjoint_partition(same, 3)] == [2, 2, 2]
    assert GI.BoolFlag.BOOL_TRUE == 1
    with pytest.raises(dgl.DGLError):
        GI.create_graph_index(3.5, True)


def test_random_choice_native():
    dgl.seed(1)
    x = dgl.random.choice(1000000, 100, replace=False)
    assert x.unique().numel() == 100 and x.max() < 1000000
    assert sorted(dgl.random.choice(100, 100, replace=False).tolist()) == list(range(100))
    assert dgl.random.choice(10, (3, 4)).shape == (3, 4)
    p = torch.tensor([0.0, 1.0, 0.0, 3.0])
    c = torch.bincount(dgl.random.choice(4, 40000, prob=p), minlength=4)
    assert c[0] == 0 and c[2] == 0 and abs(c[3].item() / 40000 - 0.75) < 0.02
    pick = dgl.random.choice(torch.tensor([10, 20, 30, 40]), 2, replace=False, prob=p)
    assert sorted(pick.tolist()) == [20, 40]
    with pytest.raises(Exception):
        dgl.random.choice(4, 3, replace=False, prob=p)
    with pytest.raises(Exception):
        dgl.random.choice(4, 5, replace=False)
    w = torch.rand(500, dtype=torch.float64)
    cnt = torch.zeros(500)
    for _ in range(1500):
        cnt[dgl.random.choice(500, 10, replace=False, prob=w)] += 1
    assert np.corrcoef(cnt.numpy(), w.numpy())[0, 1] > 0.85
    dgl.seed(5)
    a = dgl.random.choice(100, 5)
    dgl.seed(5)
    assert torch.equal(a, dgl.random.choice(100, 5))


def test_small_module_helpers(hg):
    assert dgl.init.zero_initializer((2, 3), torch.float32, "cpu", slice(0, 2)).eq(0).all()
    with pytest.raises(NotImplementedError):
        dgl.init.base_initializer((1,), torch.float32, "cpu", slice(0, 1))
    g = dgl.convert.create_from_edges("coo", (torch.tensor([0, 1]), torch.tensor([2, 0])), "u", "e", "v", 2, 3)
    assert g.num_nodes("u") == 2 and g.num_nodes("v") == 3 and g.num_edges() == 2
    g = dgl.convert.create_from_edges("csr", (torch.tensor([0, 1, 2, 2]), torch.tensor([2, 0]), torch.tensor([], dtype=torch.int64)),
                                      "n", "e", "n", 3, 3)
    assert g.edges()[1].tolist() == [2, 0] and g.is_homogeneous
    h = dgl.partition.reorder_nodes(dgl.graph(([0, 1], [1, 2])), torch.tensor([2, 0, 1]))
    assert h.edges()[0].tolist() == [2, 0] and h.edges()[1].tolist() == [0, 1] and h.ndata["orig_id"].tolist() == [1, 2, 0]
    with pytest.raises(dgl.DGLError):
        dgl.partition.reorder_nodes(dgl.graph(([0], [1])), torch.tensor([0, 0]))
    assert dgl.partition.get_peak_mem() >= 0.0
    assert dgl.core.data_dict_to_list(hg, {"a": 1, "b": 2}, fn.u_mul_v("x", "y", "m"), "u") == [1, 2]
    assert dgl.core.data_dict_to_list(hg, {"a": 1, "b": 2}, fn.u_mul_v("x", "y", "m"), "v") == [1, 2]
    rels = {("a", "r", "b"): 1, ("b", "s", "a"): 2}
    assert dgl.core.data_dict_to_list(hg, rels, fn.u_mul_e("x", "y", "m"), "e") == [1, 2]
    assert dgl.core.data_dict_to_list(hg, {"a": 5, "b": 6}, fn.copy_u("x", "m"), "u") == [5, 6]
    assert dgl.base.dgl_warning_format("hi", dgl.base.DGLWarning, "f", 1) == "DGL Warning: hi\n"
    assert "UserWarning" in dgl.base.dgl_warning_format("hi", UserWarning, "f.py", 1)
    s, d = dgl.homophily.get_long_edges(dgl.graph(([0], [1]), idtype=torch.int32))
    assert s.dtype == torch.int64 and d.tolist() == [1]


def test_pickle_states_and_sparse_ops_leftovers(hg):
    st = pickle.loads(pickle.dumps(H.HeteroPickleStates.from_index(hg._graph)))
    assert st.version == 2 and len(st.arrays) == 4
    back = st.to_index().graph
    assert back.canonical_etypes == hg.canonical_etypes and back.num_edges() == hg.num_edges()
    assert back.num_nodes("b") == hg.num_nodes("b")
    f = H.FlattenedHeteroGraph(hg, *range(9))
    assert f.induced_srctype == 0 and f.induced_dstid == 8
    from dgl_b200 import _sparse_ops as S
    assert S.get_typeid_by_target(hg._graph, 0, "u") == 0 and S.get_typeid_by_target(hg._graph, 0, "v") == 1
    assert S.get_typeid_by_target(hg._graph, 1, "e") == 1 and S.to_dgl_nd(3) == 3 and S.to_dgl_nd_for_write(4) == 4
    g = dgl.rand_graph(30, 90)
    u, v = g.edges()
    out, cw = torch.zeros(90, dtype=torch.int64), torch.zeros(3)
    S.libra_vertex_cut(3, None, None, cw, u, v, None, out, 30, 90, "x")
    assert int(cw.sum()) == 90 and set(out.tolist()) <= {0, 1, 2} and cw.max() - cw.min() <= 3
    key = torch.zeros(10, dtype=torch.int64)
    assert S.libra2dgl_build_dict(torch.tensor([5, 5, 7]), torch.tensor([7, 9, 5]), None, key, None, None, None, 0, 3, 0, 1, "x") == 3
    assert sorted(key[:3].tolist()) == [5, 7, 9]
