"""Dataset zoo, temporal GraphBolt sampler, aligned npy writer, extra nn layers."""
import numpy as np
import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.graphbolt as gb
import dgl_b200.nn as dnn
from dgl_b200 import data as D


@pytest.fixture(autouse=True)
def _tmp_download_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("DGL_DOWNLOAD_DIR", str(tmp_path))


@pytest.mark.parametrize("make,kind", [
    (lambda: D.CiteseerGraphDataset(), "node"), (lambda: D.WikiCSDataset(scale=0.1), "node"),
    (lambda: D.YelpDataset(scale=0.001), "node"), (lambda: D.KarateClubDataset(), "node"),
    (lambda: D.TUDataset("ENZYMES", scale=0.05), "graphs"), (lambda: D.GINDataset("MUTAG", self_loop=True), "graphs"),
    (lambda: D.QM9Dataset(scale=0.0005), "graphs"), (lambda: D.ZINCDataset("test", scale=0.01), "graphs"),
    (lambda: D.QM7bDataset(scale=0.01), "graphs"), (lambda: D.FakeNewsDataset("politifact", scale=0.05), "graphs"),
    (lambda: D.PPIDataset("valid"), "plain"), (lambda: D.SSTDataset("tiny", scale=1.0), "plain"),
    (lambda: D.ICEWS18Dataset(), "plain"), (lambda: D.BitcoinOTCDataset(), "plain"),
    (lambda: D.FB15k237Dataset(scale=0.01), "kg"), (lambda: D.AIFBDataset(scale=0.1), "rdf"),
    (lambda: D.FraudDataset("amazon", scale=0.05), "rdf")])
def test_dataset_zoo(make, kind):
    ds = make()
    assert len(ds) >= 1
    x = ds[0]
    if kind == "node":
        assert x.num_nodes() > 0 and "label" in x.ndata
    elif kind == "graphs":
        g, y = x
        assert g.num_nodes() > 0 and torch.is_tensor(y)
    elif kind == "kg":
        assert "etype" in x.edata and int(x.edata["train_mask"].sum()) > 0 and ds.num_rels > 0
    elif kind == "rdf":
        assert len(x.canonical_etypes) >= 3
    else:
        assert x.num_nodes() > 0


def test_karate_is_real_and_sbm_is_blocky():
    g = D.KarateClubDataset()[0]
    assert g.num_nodes() == 34 and g.num_edges() == 156 and int(g.ndata["label"].sum()) == 17
    ds = D.SBMMixtureDataset(2, 60, 2, pq=(0.5, 0.02))
    g, lab = ds[0]
    u, v = g.edges()
    same = (lab[u] == lab[v]).float().mean()
    assert same > 0.85  # p >> q: most edges stay inside a block
    bg, bl = ds.collate_fn([ds[0], ds[1]])
    assert bg.num_nodes() == 120 and bl.shape[0] == 120


def test_dataset_adapters():
    gp = D.AsGraphPredDataset(D.TUDataset("MUTAG", scale=0.2))
    assert len(gp.train_idx) + len(gp.val_idx) + len(gp.test_idx) == len(gp)
    lp = D.AsLinkPredDataset(D.CiteseerGraphDataset())
    (pu, pv), (nu, nv) = lp.val_edges
    assert nu.shape[0] == 3 * pu.shape[0] and lp.train_graph.num_edges() > 0


def test_numpy_save_aligned(tmp_path):
    a = np.random.rand(100, 7).astype(np.float32)
    p = str(tmp_path / "a.npy")
    gb.numpy_save_aligned(p, a)
    assert (np.load(p) == a).all()
    with open(p, "rb") as f:
        f.read(8)
        hl = int.from_bytes(f.read(4), "little")
    assert (12 + hl) % 4096 == 0


def test_temporal_neighbor_sampler():
    torch.manual_seed(0)
    g = dgl.rand_graph(200, 3000)
    fg = gb.from_dglgraph(g)
    fg.node_attributes = {"ts": torch.randint(0, 100, (200,))}
    fg.edge_attributes = dict(fg.edge_attributes or {})
    fg.edge_attributes["ts"] = torch.randint(0, 100, (fg.indices.shape[0],))
    items = gb.ItemSet((torch.arange(50), torch.randint(50, 100, (50,))), names=("seeds", "timestamp"))
    dp = gb.TemporalNeighborSampler(gb.ItemSampler(items, batch_size=10), fg, [3, 2], node_timestamp_attr_name="ts",
                                    edge_timestamp_attr_name="ts")
    n = 0
    for mb in dp:
        sg = mb.sampled_subgraphs[-1]  # seed layer
        deg = (sg.sampled_csc.indptr[1:] - sg.sampled_csc.indptr[:-1]).long()
        assert int(deg.max()) <= 2
        seed_ts = mb.timestamp
        src = sg.original_row_node_ids[sg.sampled_csc.indices.long()]
        owner = torch.repeat_interleave(torch.arange(10), deg)
        assert bool((fg.node_attributes["ts"][src.long()] < seed_ts[owner]).all())
        assert bool((fg.edge_attributes["ts"][sg.original_edge_ids.long()] < seed_ts[owner]).all())
        n += 1
    assert n == 5


def test_twirls_unfolding_and_grouprevres():
    g = dgl.add_self_loop(dgl.rand_graph(100, 600))
    x = torch.randn(100, 8, requires_grad=True)
    y = dnn.TWIRLSUnfoldingAndAttention(8, 0, 1.0, 4, attn_aft=2)(g, x)
    assert y.shape == (100, 8)
    m = dnn.GroupRevRes(dnn.GraphConv(4, 4), groups=2)
    out = m(g, x)
    out.sum().backward()
    assert out.shape == (100, 8) and x.grad is not None
    # definition check: Y_1 = f_1(X_2) + X_1, Y_2 = f_2(Y_1) + X_2
    x1, x2 = x.detach().chunk(2, -1)
    y1 = m.gnn_modules[0](g, x2) + x1
    y2 = m.gnn_modules[1](g, y1) + x2
    assert torch.allclose(out.detach(), torch.cat([y1, y2], -1), atol=1e-6)
    mc = dnn.GroupRevRes(dnn.GraphConv(4, 4), groups=2, checkpoint=True)
    mc(g, x).sum().backward()
