"""knn / permutation helpers / module-level helpers of ``dgl.transforms`` (reference transforms/functional.py:100,641,
3347,3371 and transforms/module.py:60,1294)."""
import networkx as nx
import numpy as np
import pytest
import scipy.sparse as sp
import torch
from scipy.sparse.csgraph import reverse_cuthill_mckee

import dgl_b200 as dgl
from dgl_b200.transforms import functional as TF
from dgl_b200.transforms import module as TM


def _bandwidth(g, perm):
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    u, v = [t.numpy() for t in g.edges()]
    return int(np.abs(inv[u] - inv[v]).max())


@pytest.mark.parametrize("make", [lambda: nx.grid_2d_graph(20, 20), lambda: nx.barabasi_albert_graph(800, 2, seed=0),
                                  lambda: nx.disjoint_union(nx.path_graph(40), nx.cycle_graph(60))])
def test_native_rcmk_matches_scipy_bandwidth(make):
    g = dgl.from_networkx(nx.convert_node_labels_to_integers(make()))
    torch.manual_seed(0)
    g = dgl.reorder_graph(g, "custom", permute_config={"nodes_perm": torch.randperm(g.num_nodes())}, store_ids=False)
    perm = TF.rcmk_perm(g)
    assert sorted(perm.tolist()) == list(range(g.num_nodes()))
    u, v = [t.numpy() for t in g.edges()]
    ref = reverse_cuthill_mckee(sp.coo_matrix((np.ones(len(u)), (u, v)), shape=(g.num_nodes(),) * 2).tocsr(), True)
    assert _bandwidth(g, perm) <= 1.1 * _bandwidth(g, ref) + 1
    assert _bandwidth(g, perm) < _bandwidth(g, np.arange(g.num_nodes()))
    r = dgl.reorder_graph(g, "rcmk")
    assert r.num_edges() == g.num_edges() and dgl.NID in r.ndata


def test_metis_perm_groups_partitions():
    g = dgl.from_networkx(nx.grid_2d_graph(12, 12))
    g = dgl.graph(g.edges(), num_nodes=144)
    perm = TF.metis_perm(g, 4)
    assert sorted(perm.tolist()) == list(range(144)) and isinstance(perm, np.ndarray)


def test_knn_function_segments_and_metrics():
    torch.manual_seed(0)
    x = torch.randn(300, 3)
    out = TF.knn(4, x, [100, 200])
    assert out.shape == (2, 1200) and out.dtype == torch.int64
    assert out[0].view(300, 4)[:, 0].tolist() == list(range(300))
    assert torch.all(out[1].view(300, 4)[:, 0] == torch.arange(300))          # nearest point of a self query is itself
    assert torch.all(out[1][:400] < 100) and torch.all(out[1][400:] >= 100)
    y = torch.randn(30, 3)
    for algo in ("kd-tree", "bruteforce", "bruteforce-blas", "nn-descent"):
        if algo == "nn-descent":
            with pytest.raises(dgl.DGLError):
                TF.knn(2, x, [100, 200], y, [10, 20], algorithm=algo)
            continue
        o = TF.knn(2, x, [100, 200], y, [10, 20], algorithm=algo)
        d = torch.cdist(y[:10], x[:100], compute_mode="donot_use_mm_for_euclid_dist")
        assert torch.equal(o[1].view(30, 2)[:10], torch.topk(d, 2, largest=False)[1])
    c = TF.knn(3, x, [300], dist="cosine")
    xn = torch.nn.functional.normalize(x, dim=1)
    assert torch.equal(c[1].view(300, 3), torch.topk(xn @ xn.t(), 3)[1]) or \
        (c[1].view(300, 3) == torch.topk(xn @ xn.t(), 3)[1]).float().mean() > 0.99
    with pytest.warns(Warning):
        assert TF.knn(500, x, [100, 200]).shape[1] == 300 * 100
    with pytest.raises(dgl.DGLError):
        TF.knn(0, x, [300])
    with pytest.raises(dgl.DGLError):
        TF.knn(2, x, [300], dist="manhattan")
    big = torch.randn(50, 40)                                                   # high-dim -> GEMM brute force
    o = TF.knn(3, big, [50])
    assert torch.equal(o[1].view(50, 3)[:, 0], torch.arange(50))


def test_pairwise_squared_distance():
    x = torch.randn(2, 7, 5)
    assert torch.allclose(TF.pairwise_squared_distance(x), torch.cdist(x, x) ** 2, atol=1e-4)


def test_is_bidirected_and_update_graph_structure():
    assert TM.is_bidirected(dgl.graph(([0, 1, 1, 2], [1, 0, 2, 1])))
    assert not TM.is_bidirected(dgl.graph(([0, 1, 1], [1, 0, 2])))
    assert not TM.is_bidirected(dgl.graph(([0, 0, 1], [1, 1, 0])))             # multiplicities differ
    g = dgl.heterograph({("a", "r", "b"): ([0, 1], [1, 0])}, {"a": 3, "b": 2})
    g.nodes["a"].data["h"] = torch.arange(3.0)
    g.edges["r"].data["w"] = torch.tensor([5.0, 6.0])
    new = TM.update_graph_structure(g, {("a", "r", "b"): ([2, 2], [0, 1])})
    assert new.num_nodes("a") == 3 and new.edges()[0].tolist() == [2, 2]
    assert new.nodes["a"].data["h"].tolist() == [0.0, 1.0, 2.0] and new.edges["r"].data["w"].tolist() == [5.0, 6.0]
    bare = TM.update_graph_structure(g, {("a", "r", "b"): ([2], [0])}, copy_edata=False)
    assert "w" not in bare.edges["r"].data
