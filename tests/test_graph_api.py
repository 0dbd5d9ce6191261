"""DGLGraph API behaviour (subset of /root/reference/tests/python/common/test_heterograph*.py, rewritten)."""
import pickle

import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.function as fn
from graph_cases import get_cases


def test_create_and_query(idtype):
    g = dgl.graph(([0, 0, 1, 2], [1, 2, 2, 3]), idtype=idtype)
    assert g.num_nodes() == 4 and g.num_edges() == 4 and g.idtype == idtype
    assert g.has_edges_between(0, 1) and not g.has_edges_between(1, 0)
    assert g.in_degrees().tolist() == [0, 1, 2, 1]
    assert g.out_degrees().tolist() == [2, 1, 1, 0]
    assert g.edge_ids(0, 2) == 1
    u, v = g.find_edges([2, 0])
    assert u.tolist() == [1, 0] and v.tolist() == [2, 1]
    assert g.successors(0).tolist() == [1, 2]
    assert g.predecessors(2).tolist() == [0, 1]
    s, d, e = g.in_edges([2, 3], form="all")
    assert sorted(e.tolist()) == [1, 2, 3]
    assert g.formats()["created"] != []
    g2 = g.formats(["csc"])
    assert g2.formats()["created"] == ["csc"] or g2.formats()["not created"] == []
    with pytest.raises(dgl.DGLError):
        g2._graphs[0].coo()


def test_csr_csc_construction(idtype):
    indptr = torch.tensor([0, 2, 3, 3, 4])
    indices = torch.tensor([1, 2, 2, 0])
    g = dgl.graph(("csr", (indptr, indices, [])), idtype=idtype)
    u, v = g.edges()
    assert u.tolist() == [0, 0, 1, 3] and v.tolist() == [1, 2, 2, 0]
    g = dgl.graph(("csc", (indptr, indices, [])), idtype=idtype)
    u, v = g.edges()
    assert v.tolist() == [0, 0, 1, 3] and u.tolist() == [1, 2, 2, 0]


def test_features_and_scope():
    g = dgl.rand_graph(10, 30)
    g.ndata["x"] = torch.ones(10, 3)
    g.edata["w"] = torch.ones(30)
    with g.local_scope():
        g.ndata["y"] = torch.zeros(10)
        g.ndata["x"] = torch.zeros(10, 3)
        assert "y" in g.ndata
    assert "y" not in g.ndata and g.ndata["x"].sum() == 30
    with pytest.raises(dgl.DGLError):
        g.ndata["bad"] = torch.zeros(9)
    g.nodes[[0, 1]].data["x"] = torch.full((2, 3), 5.0)
    assert g.ndata["x"][0, 0] == 5 and g.ndata["x"][2, 0] == 1
    assert g.node_attr_schemes()["x"] == ((3,), torch.float32)


def test_hetero_basic(idtype):
    g = get_cases(["hetero"])[0](idtype)
    assert set(g.ntypes) == {"user", "game", "developer"}
    assert g.num_nodes("user") == 3 and g.num_edges("plays") == 4
    assert g.to_canonical_etype("plays") == ("user", "plays", "game")
    sub = g["user", "plays", "game"]
    assert sub.num_edges() == 4 and sub.is_unibipartite
    g.nodes["user"].data["h"] = torch.ones(3, 2)
    g.nodes["developer"].data["h"] = torch.ones(2, 2) * 2
    g.multi_update_all({"plays": (fn.copy_u("h", "m"), fn.sum("m", "y")),
                        "develops": (fn.copy_u("h", "m"), fn.sum("m", "y"))}, "sum")
    assert g.nodes["game"].data["y"].tolist() == [[4.0, 4.0], [4.0, 4.0]]
    hg = dgl.to_homogeneous(g)
    assert hg.num_nodes() == g.num_nodes() and hg.num_edges() == g.num_edges()
    back = dgl.to_heterogeneous(hg, g.ntypes, g.etypes)
    assert back.num_edges("plays") == 4
    assert g.node_type_subgraph(["user"]).num_edges() == 3


def test_udf_paths_match_builtin():
    g = get_cases(["random"])[0]()
    g.ndata["h"] = torch.randn(g.num_nodes(), 4)
    g.edata["w"] = torch.randn(g.num_edges(), 4)
    g.update_all(fn.u_mul_e("h", "w", "m"), fn.sum("m", "a"))
    g.update_all(lambda e: {"m": e.src["h"] * e.data["w"]}, lambda n: {"b": n.mailbox["m"].sum(1)})
    assert torch.allclose(g.ndata["a"], g.ndata["b"], atol=1e-5)
    g.update_all(fn.copy_u("h", "m"), fn.max("m", "c"))
    g.update_all(lambda e: {"m": e.src["h"]}, lambda n: {"d": n.mailbox["m"].max(1)[0]})
    assert torch.allclose(g.ndata["c"], g.ndata["d"])
    g.apply_edges(fn.u_add_v("h", "h", "s"))
    g.apply_edges(lambda e: {"t": e.src["h"] + e.dst["h"]})
    assert torch.allclose(g.edata["s"], g.edata["t"])
    g.apply_nodes(lambda n: {"h2": n.data["h"] * 2})
    assert torch.allclose(g.ndata["h2"], g.ndata["h"] * 2)
    # partial message passing
    g.ndata["p"] = torch.zeros(g.num_nodes(), 4)
    g.pull([3, 5], fn.copy_u("h", "m"), fn.sum("m", "p"))
    u, _ = g.in_edges(3)
    assert torch.allclose(g.ndata["p"][3], g.ndata["h"][u].sum(0), atol=1e-5)
    g.send_and_recv([0, 1, 2], fn.copy_u("h", "m"), fn.sum("m", "q"))


def test_mutation_and_subgraph(idtype):
    g = dgl.graph(([0, 1], [1, 2]), idtype=idtype)
    g.ndata["x"] = torch.arange(3).float()
    g.add_nodes(2)
    g.add_edges([3, 4], [0, 0])
    assert g.num_nodes() == 5 and g.num_edges() == 4 and g.ndata["x"].shape[0] == 5
    g.remove_edges([0])
    assert g.num_edges() == 3
    g.remove_nodes([4])
    assert g.num_nodes() == 4 and g.num_edges() == 2
    g = dgl.rand_graph(50, 300, idtype=idtype)
    g.ndata["x"] = torch.arange(50).float()
    sg = g.subgraph([1, 5, 7, 9, 20])
    assert sg.ndata[dgl.NID].tolist() == [1, 5, 7, 9, 20]
    assert torch.equal(sg.ndata["x"], torch.tensor([1.0, 5, 7, 9, 20]))
    u, v = sg.edges()
    pu, pv = sg.ndata[dgl.NID][u.long()], sg.ndata[dgl.NID][v.long()]
    assert bool(g.has_edges_between(pu, pv).all())
    esg = g.edge_subgraph([0, 1, 2])
    assert esg.num_edges() == 3
    isg = dgl.in_subgraph(g, [0, 1])
    assert isg.num_edges() == int(g.in_degrees([0, 1]).sum())
    ksg, inv = dgl.khop_in_subgraph(g, 0, 2)
    assert ksg.ndata[dgl.NID][inv.long()].tolist() == [0]


def test_batch_readout():
    gs = [dgl.rand_graph(n, 3 * n) for n in (4, 6, 3)]
    for g in gs:
        g.ndata["h"] = torch.randn(g.num_nodes(), 2)
        g.edata["e"] = torch.randn(g.num_edges(), 2)
    bg = dgl.batch(gs)
    assert bg.batch_size == 3 and bg.batch_num_nodes().tolist() == [4, 6, 3]
    s = dgl.sum_nodes(bg, "h")
    assert torch.allclose(s, torch.stack([g.ndata["h"].sum(0) for g in gs]), atol=1e-5)
    m = dgl.max_edges(bg, "e")
    assert torch.allclose(m, torch.stack([g.edata["e"].max(0)[0] for g in gs]))
    parts = dgl.unbatch(bg)
    assert [p.num_nodes() for p in parts] == [4, 6, 3]
    assert torch.equal(parts[1].edges()[0], gs[1].edges()[0])
    assert dgl.slice_batch(bg, 1).num_edges() == 18
    assert dgl.broadcast_nodes(bg, torch.arange(3.0).view(3, 1)).shape == (13, 1)
    tk, idx = dgl.topk_nodes(bg, "h", 2)
    assert tk.shape == (3, 2, 2)


def test_transforms(idtype):
    g = dgl.graph(([0, 1, 1], [1, 2, 2]), num_nodes=4, idtype=idtype)
    assert dgl.add_self_loop(g).num_edges() == 7
    assert dgl.remove_self_loop(dgl.add_self_loop(g)).num_edges() == 3
    sg = dgl.to_simple(g)
    assert sg.num_edges() == 2 and sg.edata["count"].tolist() == [1, 2]
    bg = dgl.to_bidirected(g)
    assert bg.num_edges() == 4
    rg = dgl.reverse(g)
    assert rg.edges()[0].tolist() == [1, 2, 2]
    arg = dgl.add_reverse_edges(g)
    assert arg.num_edges() == 6
    lg = dgl.line_graph(g)
    assert lg.num_nodes() == 3
    assert dgl.khop_graph(g, 2).num_edges() == 2
    cg = dgl.compact_graphs(g)
    assert cg.num_nodes() == 3
    pe = dgl.random_walk_pe(dgl.add_self_loop(g), 3)
    assert pe.shape == (4, 3)
    t = dgl.Compose([dgl.AddSelfLoop(), dgl.GCNNorm()])
    out = t(dgl.graph(([0, 1], [1, 0])))
    assert "w" in out.edata
    kg = dgl.knn_graph(torch.randn(20, 3), 4)
    assert kg.num_edges() == 80


def test_to_block_semantics(idtype):
    g = dgl.graph(([5, 6, 7, 5], [0, 0, 1, 2]), num_nodes=8, idtype=idtype)
    g.ndata["x"] = torch.arange(8).float()
    b = dgl.to_block(g, torch.tensor([0, 1, 2], dtype=idtype))
    assert b.num_dst_nodes() == 3 and b.num_src_nodes() == 6
    assert b.srcdata[dgl.NID].tolist() == [0, 1, 2, 5, 6, 7]
    assert b.dstdata[dgl.NID].tolist() == [0, 1, 2]
    assert torch.equal(b.srcdata["x"], torch.tensor([0.0, 1, 2, 5, 6, 7]))
    u, v = b.edges()
    assert u.tolist() == [3, 4, 5, 3] and v.tolist() == [0, 0, 1, 2]


def test_pickle_and_clone():
    g = dgl.rand_graph(20, 50)
    g.ndata["x"] = torch.randn(20, 2)
    g2 = pickle.loads(pickle.dumps(g))
    assert torch.equal(g2.edges()[0], g.edges()[0]) and torch.equal(g2.ndata["x"], g.ndata["x"])
    g3 = g.int()
    assert g3.idtype == torch.int32 and g3.long().idtype == torch.int64
