"""Every model family of the zoo trains for a few steps on a tiny graph (loss finite and decreasing gradients flow)."""
import pytest
import torch
import torch.nn.functional as F

import dgl_b200 as dgl
from dgl_b200 import models as M


def _node_data(n=120, e=900, f=12, c=4, seed=0):
    torch.manual_seed(seed)
    g = dgl.add_self_loop(dgl.to_bidirected(dgl.rand_graph(n, e)))
    return g, torch.randn(n, f), torch.randint(0, c, (n,))


def _fit(model, fwd, target, steps=8, lr=1e-2):
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    for _ in range(steps):
        loss = F.cross_entropy(fwd(), target)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(l == l for l in losses) and losses[-1] < losses[0]
    return losses


@pytest.mark.parametrize("name", ["SGC", "APPNP", "GATv2", "ChebNet", "TAGCN", "GCNII", "JKNet"])
def test_node_models(name):
    g, x, y = _node_data()
    ctor = {"SGC": lambda: M.SGC(12, 4), "APPNP": lambda: M.APPNP(12, 16, 4, dropout=0.0), "GATv2": lambda: M.GATv2(12, 8, 4),
            "ChebNet": lambda: M.ChebNet(12, 16, 4), "TAGCN": lambda: M.TAGCN(12, 16, 4),
            "GCNII": lambda: M.GCNII(12, 16, 4, num_layers=4, dropout=0.0), "JKNet": lambda: M.JKNet(12, 16, 4)}[name]
    m = ctor()
    _fit(m, lambda: m(g, x), y)


def test_gin_graph_classification():
    torch.manual_seed(0)
    graphs = [dgl.add_self_loop(dgl.rand_graph(10 + i % 5, 40)) for i in range(16)]
    bg = dgl.batch(graphs)
    x = torch.randn(bg.num_nodes(), 6)
    y = torch.randint(0, 3, (16,))
    m = M.GIN(6, 16, 3, num_layers=3)
    _fit(m, lambda: m(bg, x), y)


def test_link_predictor_and_dgi():
    g, x, _ = _node_data()
    u, v = g.edges()
    pos = dgl.graph((u[:200], v[:200]), num_nodes=g.num_nodes())
    nu, nv = dgl.sampling.global_uniform_negative_sampling(g, 200)
    neg = dgl.graph((nu, nv), num_nodes=g.num_nodes())
    for scorer in ("dot", "cos"):
        m = M.LinkPredictor(12, 16, scorer=scorer)
        opt = torch.optim.Adam(m.parameters(), lr=1e-2)
        first = last = None
        for _ in range(8):
            loss = m.loss(*m(g, x, pos, neg))
            opt.zero_grad(); loss.backward(); opt.step()
            first = float(loss) if first is None else first
            last = float(loss)
        assert last == last and last <= first
    dgi = M.DGI(12, 16)
    opt = torch.optim.Adam(dgi.parameters(), lr=1e-2)
    for _ in range(5):
        loss = dgi(g, x)
        opt.zero_grad(); loss.backward(); opt.step()
    assert float(loss) == float(loss) and dgi.embed(g, x).shape == (g.num_nodes(), 16)


def test_hetero_models():
    torch.manual_seed(0)
    hg = dgl.heterograph({("user", "buys", "item"): (torch.randint(0, 30, (200,)), torch.randint(0, 20, (200,))),
                          ("item", "rev_buys", "user"): (torch.randint(0, 20, (200,)), torch.randint(0, 30, (200,)))})
    feats = {"user": torch.randn(30, 8), "item": torch.randn(20, 6)}
    y = torch.randint(0, 3, (30,))
    m = M.HeteroRGCN({"user": 8, "item": 6}, 16, 3, hg.etypes)
    _fit(m, lambda: m(hg, feats)["user"], y)
    g = dgl.to_homogeneous(hg)
    x = torch.randn(g.num_nodes(), 8)
    hgt = M.HGT(8, 16, 3, len(hg.ntypes), len(hg.etypes), heads=4)
    yy = torch.randint(0, 3, (g.num_nodes(),))
    _fit(hgt, lambda: hgt(g, x, g.ndata[dgl.NTYPE], g.edata[dgl.ETYPE]), yy)


def test_label_propagation_model():
    g, _, y = _node_data()
    mask = torch.zeros(g.num_nodes(), dtype=torch.bool)
    mask[:60] = True
    out = M.LabelPropagationModel(k=5)(g, y, mask)
    assert out.shape[0] == g.num_nodes()
