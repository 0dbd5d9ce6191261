"""``dgl.data`` name-dispatching bases, loaders, generators, archives and structural-shift splits
(reference python/dgl/data/{__init__,gnn_benchmark,geom_gcn,heterophilous_graphs,knowledge_graph,rdf,sbm,
tensor_serialize,heterograph_serialize,adapter,utils}.py)."""
import argparse
import os

import networkx as nx
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import dgl_b200 as dgl
from dgl_b200 import data as D
from dgl_b200.data import utils as DU


def test_named_bases_and_loaders():
    p = argparse.ArgumentParser()
    D.register_data_args(p)
    ds = D.load_data(p.parse_args(["--dataset", "cora"]))
    assert ds[0].num_nodes() == 2708
    with pytest.raises(ValueError):
        D.load_data(p.parse_args(["--dataset", "nope"]))
    assert type(D.GeomGCNDataset("cornell")).__name__ == "CornellDataset"
    assert type(D.GNNBenchmarkDataset("coauthor_cs", scale=0.05)).__name__ == "CoauthorCSDataset"
    assert type(D.HeterophilousGraphDataset("minesweeper", scale=0.05)).__name__ == "MinesweeperDataset"
    assert D.KnowledgeGraphDataset("FB15k-237", scale=0.01)[0].num_edges() > 0
    assert len(D.RDFGraphDataset("aifb", scale=0.05)[0].ntypes) > 1
    with pytest.raises(dgl.DGLError):
        D.GeomGCNDataset("unknown")
    assert D.load_citeseer()[0].num_nodes() == 3327


def test_sbm_generator_statistics():
    rng = np.random.RandomState(0)
    n_blocks, size, p, q = 4, 200, 16.0, 2.0
    A = D.sbm.sbm(n_blocks, size, p, q, rng)     # dgl.data.sbm is the module, as in the reference
    n = n_blocks * size
    assert A.shape == (n, n) and abs(A - A.T).nnz == 0
    comm = np.arange(n) // size
    same = comm[A.row] == comm[A.col]
    # off-diagonal intra-block pairs are drawn once per block with prob p/n, inter-block with q/n
    intra = same.sum() / (n_blocks * size * size)
    inter = (~same).sum() / (n_blocks * (n_blocks - 1) * size * size)
    assert abs(intra - p / n) / (p / n) < 0.15 and abs(inter - q / n) / (q / n) < 0.25


def test_build_knowledge_graph_and_negative_sample():
    g, parts = D.build_knowledge_graph(5, 2, [[0, 0, 1]], [[1, 1, 2]], [[2, 0, 3]])
    assert g.num_edges() == 6 and g.edata["etype"].tolist() == [0, 1, 0, 2, 3, 2]
    assert g.edata["train_edge_mask"].tolist() == [True, False, False, True, False, False]
    assert g.edges()[0].tolist() == [0, 1, 2, 1, 2, 3] and len(parts) == 3
    g2, _ = D.build_knowledge_graph(5, 2, [[0, 0, 1]], [[1, 1, 2]], [[2, 0, 3]], reverse=False)
    assert g2.num_edges() == 3
    gr = dgl.rand_graph(60, 300)
    neg = D.negative_sample(gr, 40)
    assert neg.shape == (2, 40) and not np.any(neg[0] == neg[1])
    assert not gr.has_edges_between(torch.as_tensor(neg[0]), torch.as_tensor(neg[1])).any()
    assert len({(a, b) for a, b in neg.T.tolist()}) == 40
    assert D.eliminate_self_loops((sp.eye(3) + sp.coo_matrix(([1.0], ([0], [1])), shape=(3, 3))).tocsr()).nnz == 1


def test_archives(tmp_path):
    t = os.path.join(tmp_path, "sub", "t.bin")
    assert D.save_tensors(t, {"a": torch.arange(3), "b": torch.ones(2, 2)})
    back = D.load_tensors(t)
    assert back["a"].tolist() == [0, 1, 2] and back["b"].shape == (2, 2)
    with pytest.raises(dgl.DGLError):
        D.save_tensors(t, {1: torch.zeros(1)})
    with pytest.raises(dgl.DGLError):
        D.save_tensors(t, {"a": [1, 2]})
    hg = dgl.heterograph({("a", "r", "b"): ([0, 1], [1, 0])})
    gp = os.path.join(tmp_path, "g.bin")
    D.save_heterographs(gp, [hg], {"y": torch.tensor([3])})
    gs, labels = D.load_graph_v2(gp)
    assert gs[0].canonical_etypes == hg.canonical_etypes and labels["y"].tolist() == [3]
    assert D.load_labels_v2(gp)["y"].tolist() == [3]
    meta = D.StorageMetaData.load(gp)
    assert meta.num_graphs == 1
    assert D.HeteroGraphData.create(hg).get_graph().num_edges() == 2
    assert D.is_local_path("/tmp/x") and not D.is_local_path("hdfs://x")
    with pytest.raises(dgl.DGLError):
        D.check_local_file_exists(os.path.join(tmp_path, "missing.bin"))
    with pytest.raises(dgl.DGLError):
        D.load_tensors(gp)


def test_structural_property_values_match_networkx():
    G = nx.barabasi_albert_graph(200, 3, seed=0)
    g = dgl.from_networkx(G)
    adj = DU._undirected_simple(g)
    pr = DU._pagerank(adj).numpy()
    assert np.abs(pr - np.array(list(nx.pagerank(G, tol=1e-10).values()))).max() < 1e-4
    cl = DU._clustering(adj).numpy()
    assert np.abs(cl - np.array(list(nx.clustering(G).values()))).max() < 1e-6
    pers = {i: 0.0 for i in range(200)}
    pers[int(pr.argmax())] = 1.0
    ppr = DU._pagerank(adj, torch.tensor(list(pers.values()))).numpy()
    assert np.abs(ppr - np.array(list(nx.pagerank(G, personalization=pers, tol=1e-10).values()))).max() < 1e-4


@pytest.mark.parametrize("prop", ["popularity", "locality", "density"])
def test_add_node_property_split(prop):
    g = dgl.from_networkx(nx.barabasi_albert_graph(300, 3, seed=1))
    ratios = [0.3, 0.1, 0.1, 0.3, 0.2]
    DU.add_node_property_split([g], ratios, prop, random_seed=0)
    names = ["in_train_mask", "in_valid_mask", "in_test_mask", "out_valid_mask", "out_test_mask"]
    masks = torch.stack([g.ndata[k] for k in names])
    assert masks.sum(1).tolist() == [90, 30, 30, 90, 60] and int(masks.sum(0).max()) == 1 and int(masks.sum()) == 300
    vals = -DU._node_property(g, prop, True)          # the raw property (larger = more popular / local / dense)
    # ascending=True keeps high-property nodes in distribution and pushes low-property nodes out
    assert vals[masks[:3].any(0).numpy()].mean() >= vals[masks[4].numpy()].mean()
    m1 = DU.mask_nodes_by_property(np.arange(10.0), [0.2, 0.2, 0.2, 0.2, 0.2], random_seed=3)
    m2 = DU.mask_nodes_by_property(np.arange(10.0), [0.2, 0.2, 0.2, 0.2, 0.2], random_seed=3)
    assert all(torch.equal(m1[k], m2[k]) for k in names)
    assert torch.nonzero(m1["out_test_mask"]).view(-1).tolist() == [8, 9]
    with pytest.raises(AssertionError):
        DU.add_node_property_split([g], ratios, "bogus")


def test_misc_utils(tmp_path):
    f = os.path.join(tmp_path, "a.txt")
    with open(f, "w") as h:
        h.write("1,2\n3,4\n")
    assert DU.loadtxt(f, ",").tolist() == [[1, 2], [3, 4]]
    import hashlib
    assert DU.check_sha1(f, hashlib.sha1(open(f, "rb").read()).hexdigest()) and not DU.check_sha1(f, "0" * 40)
    assert DU.idx2mask([1, 3], 5).tolist() == [0, 1, 0, 1, 0]
    with pytest.warns(UserWarning):
        DU.deprecate_function("old", "new")
