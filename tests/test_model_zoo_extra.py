"""The second half of the model zoo (dgl_b200/models/zoo_extra.py): every family trains on a small planted problem and
must actually learn it (accuracy / loss thresholds), and the custom layers are checked against dense formulas."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import dgl_b200 as dgl
from dgl_b200.models import zoo_extra as Z


def planted(n_blocks=3, size=40, p_in=0.25, p_out=0.01, feat=8, noise=1.0, seed=0):
    """SBM graph whose communities are the classes; features = noisy class means."""
    rng = np.random.RandomState(seed)
    n = n_blocks * size
    y = np.repeat(np.arange(n_blocks), size)
    prob = np.where(y[:, None] == y[None, :], p_in, p_out)
    adj = np.triu(rng.rand(n, n) < prob, 1)
    u, v = np.nonzero(adj | adj.T)
    g = dgl.add_self_loop(dgl.graph((torch.tensor(u), torch.tensor(v)), num_nodes=n))
    means = rng.randn(n_blocks, feat)
    x = torch.tensor(means[y] + noise * rng.randn(n, feat), dtype=torch.float32)
    y = torch.tensor(y)
    train = torch.zeros(n, dtype=torch.bool)
    train[torch.tensor(rng.permutation(n)[: n // 4])] = True
    return g, x, y, train


def fit_nodes(model, fwd, y, train, steps=60, lr=2e-2, extra_loss=None):
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=1e-4)
    for _ in range(steps):
        model.train()
        out = fwd()
        logits = out[0] if isinstance(out, list) else out
        loss = F.cross_entropy(logits[train], y[train])
        if extra_loss is not None:
            loss = loss + extra_loss(out)
        opt.zero_grad()
        loss.backward()
        opt.step()
    model.eval()
    with torch.no_grad():
        out = fwd()
        logits = out[0] if isinstance(out, list) else out
    return float((logits[~train].argmax(1) == y[~train]).float().mean())


@pytest.mark.parametrize("name", ["DAGNN", "MixHop", "ARMA", "DeeperGCN", "MoNet", "GRAND"])
def test_node_classifiers_learn_planted_communities(name):
    torch.manual_seed(0)
    g, x, y, train = planted()
    make = {"DAGNN": lambda: Z.DAGNN(8, 32, 3, k=5), "MixHop": lambda: Z.MixHop(8, 16, 3),
            "ARMA": lambda: Z.ARMA(8, 16, 3), "DeeperGCN": lambda: Z.DeeperGCN(8, 32, 3, num_layers=3),
            "MoNet": lambda: Z.MoNet(8, 16, 3), "GRAND": lambda: Z.GRAND(8, 32, 3, order=3, drop_node=0.3, dropout=0.2)}
    m = make[name]()
    extra = (lambda outs: Z.GRAND.consistency_loss(outs)) if name == "GRAND" else None
    acc = fit_nodes(m, lambda: m(g, x), y, train, extra_loss=extra)
    assert acc > 0.85, (name, acc)


def test_sign_precompute_matches_dense_and_learns():
    torch.manual_seed(0)
    g, x, y, train = planted()
    hops = Z.SIGN.precompute(g, x, num_hops=2)
    a = torch.zeros(g.num_nodes(), g.num_nodes())
    a[g.edges()[1].long(), g.edges()[0].long()] = 1.0
    d = a.sum(1).clamp(min=1).pow(-0.5)
    an = d[:, None] * a * d[None, :]
    assert torch.allclose(hops[1], an @ x, atol=1e-5) and torch.allclose(hops[2], an @ an @ x, atol=1e-4)
    m = Z.SIGN(8, 32, 3, num_hops=2)
    assert fit_nodes(m, lambda: m(hops), y, train) > 0.85


def test_correct_and_smooth_improves_a_weak_predictor():
    torch.manual_seed(0)
    g, x, y, train = planted(noise=2.0, seed=1)
    mlp = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    base = fit_nodes(mlp, lambda: mlp(x), y, train, steps=80)
    with torch.no_grad():
        soft = torch.softmax(mlp(x), -1)
    cs = Z.CorrectAndSmooth(num_correction_layers=20, num_smoothing_layers=20)
    out = cs.smooth(g, cs.correct(g, soft, y[train], train), y[train], train)
    acc = float((out[~train].argmax(1) == y[~train]).float().mean())
    assert acc > base + 0.1 and acc > 0.8, (base, acc)
    fixed = Z.CorrectAndSmooth(autoscale=False, scale=1.0)
    out2 = fixed.smooth(g, fixed.correct(g, soft, y[train], train), y[train], train)
    assert float((out2[~train].argmax(1) == y[~train]).float().mean()) > base


def test_gae_and_vgae_reconstruct_edges():
    torch.manual_seed(0)
    g, x, y, _ = planted()
    u, v = g.edges()
    for variational in (False, True):
        m = Z.GAE(8, 32, 16, variational=variational)
        opt = torch.optim.Adam(m.parameters(), lr=2e-2)
        for _ in range(80):
            m.train()
            z = m(g, x)
            neg = (torch.randint(0, g.num_nodes(), (len(u),)), torch.randint(0, g.num_nodes(), (len(u),)))
            loss = m.loss(z, (u, v), neg)
            opt.zero_grad()
            loss.backward()
            opt.step()
        m.eval()
        with torch.no_grad():
            z = m(g, x)
            neg = (torch.randint(0, g.num_nodes(), (2000,)), torch.randint(0, g.num_nodes(), (2000,)))
            auc_like = float((m.decode(z, u, v).mean() > m.decode(z, *neg).mean()))
            pos_s, neg_s = m.decode(z, u, v), m.decode(z, *neg)
        rank = float((pos_s[:, None] > neg_s[None, :200]).float().mean())
        assert auc_like == 1.0 and rank > 0.8, (variational, rank)


def test_grace_embeddings_separate_communities():
    torch.manual_seed(0)
    g, x, y, train = planted()
    m = Z.GRACE(8, 32, 32)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    first = last = None
    for _ in range(60):
        (g1, x1), (g2, x2) = Z.GRACE.augment(g, x), Z.GRACE.augment(g, x)
        loss = m.loss(m(g1, x1), m(g2, x2))
        opt.zero_grad()
        loss.backward()
        opt.step()
        first = float(loss) if first is None else first
        last = float(loss)
    assert last < first
    with torch.no_grad():
        z = F.normalize(m(g, x))
    # linear probe on the frozen embeddings
    probe = torch.nn.Linear(32, 3)
    assert fit_nodes(probe, lambda: probe(z), y, train, steps=100) > 0.85


def test_han_on_metapath_graphs():
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    n_paper, n_author, n_field = 90, 40, 6
    y = np.repeat(np.arange(3), 30)
    pa_u, pa_v = [], []
    for p in range(n_paper):                       # authors mostly write inside one class
        for a in rng.choice(np.arange(y[p] * 13, y[p] * 13 + 13), 2, replace=False):
            pa_u.append(p)
            pa_v.append(a)
    pf_u = np.arange(n_paper)
    pf_v = y * 2 + rng.randint(0, 2, n_paper)
    hg = dgl.heterograph({("paper", "pa", "author"): (torch.tensor(pa_u), torch.tensor(pa_v)),
                          ("author", "ap", "paper"): (torch.tensor(pa_v), torch.tensor(pa_u)),
                          ("paper", "pf", "field"): (torch.tensor(pf_u), torch.tensor(pf_v)),
                          ("field", "fp", "paper"): (torch.tensor(pf_v), torch.tensor(pf_u))},
                         {"paper": n_paper, "author": n_author, "field": n_field})
    graphs = Z.HAN.metapath_graphs(hg, [["pa", "ap"], ["pf", "fp"]])
    assert all(gr.num_nodes() == n_paper for gr in graphs)
    x = torch.randn(n_paper, 8)
    yt = torch.tensor(y)
    train = torch.zeros(n_paper, dtype=torch.bool)
    train[::3] = True
    m = Z.HAN(2, 8, 8, 3, num_heads=2)
    assert fit_nodes(m, lambda: m(graphs, x), yt, train, steps=80) > 0.8


def test_gnn_film_and_compgcn():
    torch.manual_seed(0)
    g, x, y, train = planted()
    u, v = g.edges()
    same = (y[u.long()] == y[v.long()])
    # relation 0 = intra-community edges, relation 1 = the rest: a typed view of the same graph
    hg = dgl.heterograph({("n", "intra", "n"): (u[same], v[same]), ("n", "inter", "n"): (u[~same], v[~same])},
                         {"n": g.num_nodes()})
    m = Z.GNNFiLM(hg.etypes, 8, 16, 3)
    assert fit_nodes(m, lambda: m(hg, {"n": x})["n"], y, train) > 0.85
    etype = torch.cat([(~same).long(), (~same).long() + 2])
    both = dgl.graph((torch.cat([u, v]), torch.cat([v, u])), num_nodes=g.num_nodes())
    c = Z.CompGCN(g.num_nodes(), 2, 16, 3, comp="sub", in_size=8)
    assert fit_nodes(c, lambda: c(both, etype, x), y, train, steps=80) > 0.85
    emb_only = Z.CompGCN(g.num_nodes(), 2, 16, 3, comp="mult")                  # featureless (knowledge-graph) mode
    out = emb_only(both, etype)
    out.sum().backward()
    assert out.shape == (g.num_nodes(), 3) and emb_only.node_emb.grad is not None and emb_only.rel_emb.grad is not None
    with pytest.raises(Exception):
        Z.STGCN(1, 4, 5, window=6)


def test_gcmc_rating_prediction():
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    nu, ni, k = 40, 30, 3
    pu, pi = rng.randn(nu, 2), rng.randn(ni, 2)
    score = pu @ pi.T
    cut = np.quantile(score, [1 / 3, 2 / 3])
    rating = 1 + (score > cut[0]).astype(int) + (score > cut[1]).astype(int)            # 1..3
    obs = rng.rand(nu, ni) < 0.5
    uu, ii = np.nonzero(obs)
    edges = {}
    for r in range(1, k + 1):
        sel = rating[uu, ii] == r
        edges[("user", str(r), "item")] = (torch.tensor(uu[sel]), torch.tensor(ii[sel]))
        edges[("item", f"rev-{r}", "user")] = (torch.tensor(ii[sel]), torch.tensor(uu[sel]))
    g = dgl.heterograph(edges, {"user": nu, "item": ni})
    m = Z.GCMC(k, nu, ni, agg_size=30, out_size=16)
    uf, itf = torch.eye(nu), torch.eye(ni)
    tu, ti = torch.tensor(uu), torch.tensor(ii)
    target = torch.tensor(rating[uu, ii] - 1)
    opt = torch.optim.Adam(m.parameters(), lr=2e-2)
    for _ in range(120):
        loss = F.cross_entropy(m(g, uf, itf, tu, ti), target)
        opt.zero_grad()
        loss.backward()
        opt.step()
    hu, hi = np.nonzero(~obs)
    with torch.no_grad():
        pred = m(g, uf, itf, torch.tensor(hu), torch.tensor(hi)).argmax(1) + 1
    held_out = float((pred == torch.tensor(rating[hu, hi])).float().mean())
    assert held_out > 0.5, held_out                               # chance is 1/3


def ring_vs_star(num=24, seed=0):
    rng = np.random.RandomState(seed)
    graphs, labels = [], []
    for i in range(num):
        n = int(rng.randint(8, 14))
        if i % 2 == 0:
            u = np.arange(n)
            v = (u + 1) % n
        else:
            u = np.zeros(n - 1, dtype=np.int64)
            v = np.arange(1, n)
        g = dgl.add_self_loop(dgl.to_bidirected(dgl.graph((torch.tensor(u), torch.tensor(v)), num_nodes=n)))
        graphs.append(g)
        labels.append(i % 2)
    bg = dgl.batch(graphs)
    deg = bg.in_degrees().float()
    x = torch.stack([deg, torch.ones_like(deg)], 1)
    return graphs, bg, x, torch.tensor(labels)


@pytest.mark.parametrize("name", ["DGCNN", "SAGPool", "MPNN"])
def test_graph_classifiers_tell_rings_from_stars(name):
    torch.manual_seed(0)
    graphs, bg, x, y = ring_vs_star()
    if name == "DGCNN":
        m = Z.DGCNN(2, hid_size=8, num_layers=2, k=8)
        fwd = lambda: m(bg, x)  # noqa: E731
    elif name == "SAGPool":
        m = Z.SAGPoolNet(2, 16, 2, num_layers=2)
        fwd = lambda: m(bg, x)  # noqa: E731
    else:
        m = Z.MPNN(2, 1, hid_size=8, out_size=2, steps=2)
        e = torch.ones(bg.num_edges(), 1)
        fwd = lambda: m(bg, x, e)  # noqa: E731
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(60):
        loss = F.cross_entropy(fwd(), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    with torch.no_grad():
        assert float((fwd().argmax(1) == y).float().mean()) > 0.9


def test_sagpool_keeps_top_fraction_per_graph():
    torch.manual_seed(0)
    graphs, bg, x, _ = ring_vs_star(6)
    x = torch.randn_like(x)                           # distinct scores: no ties inside a graph
    pool = Z.SAGPoolNet.Pool(2, ratio=0.5)
    sub, h = pool(bg, x)
    want = [int(np.ceil(g.num_nodes() * 0.5)) for g in graphs]
    assert sub.batch_num_nodes().tolist() == want and h.shape[0] == sum(want)
    assert int(sub.batch_num_edges().sum()) == sub.num_edges()
    with torch.no_grad():
        s = pool.score(bg, x).squeeze(-1)
    off = 0
    for g, k in zip(graphs, want):                # the kept nodes of every graph are its k best-scored ones
        seg = s[off: off + g.num_nodes()]
        kept = sub.ndata[dgl.NID][sum(want[: graphs.index(g)]): sum(want[: graphs.index(g)]) + k] - off
        assert set(kept.tolist()) == set(torch.topk(seg, k).indices.tolist())
        off += g.num_nodes()


def test_drnl_labels():
    # path 0-1-2-3-4 with target pair (0, 4): all nodes lie on the single path between the endpoints
    g = dgl.to_bidirected(dgl.graph((torch.tensor([0, 1, 2, 3]), torch.tensor([1, 2, 3, 4])), num_nodes=6))
    z = Z.drnl_node_labeling(g, 0, 4)
    # ds = [0,1,2,3,-], dt = [-,3,2,1,0]; d = 4 for the inner nodes -> 1 + min + 2 * (2 + 0 - 1)
    assert z.tolist() == [1, 4, 5, 4, 1, 0]
    tri = dgl.to_bidirected(dgl.graph((torch.tensor([0, 0, 1]), torch.tensor([1, 2, 2])), num_nodes=3))
    assert Z.drnl_node_labeling(tri, 0, 1).tolist() == [1, 1, 2]                 # common neighbour: ds = dt = 1


def test_edgeconv_point_clouds():
    torch.manual_seed(0)
    rng = np.random.RandomState(0)

    def cloud(kind):
        t = rng.rand(24) * 2 * np.pi
        if kind == 0:
            pts = np.stack([np.cos(t), np.sin(t), 0.05 * rng.randn(24)], 1)          # circle
        else:
            pts = np.stack([t / np.pi - 1, 0.05 * rng.randn(24), 0.05 * rng.randn(24)], 1)   # segment
        return pts
    y = torch.tensor([i % 2 for i in range(20)])
    x = torch.tensor(np.stack([cloud(int(k)) for k in y]), dtype=torch.float32)
    m = Z.EdgeConvNet(3, (16, 16), 2, k=4)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(40):
        loss = F.cross_entropy(m(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    with torch.no_grad():
        assert float((m(x).argmax(1) == y).float().mean()) > 0.9


def test_lgnn_community_detection():
    torch.manual_seed(0)
    g, _, y, _ = planted(n_blocks=2, size=30, p_in=0.3, p_out=0.02)
    g = dgl.remove_self_loop(g)
    lg = dgl.line_graph(g, backtracking=False)
    m = Z.LGNN(16, 2, num_layers=3, radius=2)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(150):
        out = m(g, lg)
        loss = torch.minimum(F.cross_entropy(out, y), F.cross_entropy(out, 1 - y))      # labels up to permutation
        opt.zero_grad()
        loss.backward()
        opt.step()
    m.eval()
    with torch.no_grad():
        acc = float((m(g, lg).argmax(1) == y).float().mean())
    assert max(acc, 1 - acc) > 0.9


def test_stgcn_forecasts_a_diffusing_signal():
    torch.manual_seed(0)
    n, window = 12, 10
    ring = dgl.add_self_loop(dgl.to_bidirected(dgl.graph((torch.arange(n), (torch.arange(n) + 1) % n))))
    t = torch.arange(200).float()
    series = torch.stack([torch.sin(0.2 * t + 0.5 * i) for i in range(n)], 1)             # travelling wave
    xs = torch.stack([series[i: i + window] for i in range(150)]).unsqueeze(1)            # [B, 1, T, N]
    ys = torch.stack([series[i + window] for i in range(150)])
    m = Z.STGCN(1, 8, n, window)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(80):
        loss = F.mse_loss(m(ring, xs[:120]).squeeze(1), ys[:120])
        opt.zero_grad()
        loss.backward()
        opt.step()
    m.eval()
    with torch.no_grad():
        test = float(F.mse_loss(m(ring, xs[120:]).squeeze(1), ys[120:]))
    assert test < 0.1 * float(ys.var()), test


def test_node2vec_and_pinsage_models():
    torch.manual_seed(0)
    g, x, y, train = planted(p_in=0.3)
    g = dgl.remove_self_loop(g)
    m = Z.Node2VecModel(g, emb_size=16, walk_length=8, window_size=4, num_walks=2)
    opt = torch.optim.Adam(m.parameters(), lr=5e-2)
    first = last = None
    for _ in range(40):
        pos, neg = m.sample(torch.arange(g.num_nodes()))
        loss = m.loss(pos, neg)
        opt.zero_grad()
        loss.backward()
        opt.step()
        first = float(loss) if first is None else first
        last = float(loss)
    assert last < first
    z = F.normalize(m().detach())
    probe = torch.nn.Linear(16, 3)
    assert fit_nodes(probe, lambda: probe(z), y, train, steps=100) > 0.8
    # PinSAGE layer: importance-weighted mean equals the dense formula
    block = dgl.to_block(dgl.graph((torch.tensor([0, 1, 2, 2]), torch.tensor([0, 0, 1, 0])), num_nodes=3), torch.tensor([0, 1]))
    w = torch.tensor([1.0, 2.0, 3.0, 1.0])
    layer = Z.PinSAGEModel.Layer(4, 4, 4)
    hs = torch.randn(block.num_src_nodes(), 4)
    out = layer(block, hs, hs[:2], w)
    u, v = block.edges()
    q = F.relu(layer.q(hs))
    agg = torch.zeros(2, 4).index_add_(0, v.long(), q[u.long()] * w[:, None]) / torch.zeros(2).index_add_(0, v.long(), w)[:, None]
    want = F.relu(layer.w(torch.cat([agg, hs[:2]], 1)))
    assert torch.allclose(out, want / want.norm(dim=1, keepdim=True).clamp(min=1e-6), atol=1e-5)
    assert float(Z.PinSAGEModel.margin_loss(out, out, -out)) == 0.0
