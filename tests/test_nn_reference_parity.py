"""Numerical parity of ``dgl_b200.nn`` with the UNMODIFIED reference (``baseline/_ref/dgl``) used as an oracle.

For every layer both implementations are built with the same constructor arguments, the reference's ``state_dict`` is
loaded into ours (parameter names are part of the compatibility contract), both run on the same random graph and the
outputs and input gradients must agree. Skipped when the reference build is not present."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
if not os.path.isdir(os.path.join(REF, "dgl")):
    pytest.skip("reference build not installed (baseline/_ref)", allow_module_level=True)
sys.path.insert(0, REF)
os.environ.setdefault("DGLBACKEND", "pytorch")
try:
    import dgl as rdgl
    import dgl.nn as rnn
except Exception as e:  # noqa: BLE001
    pytest.skip(f"reference not importable: {e}", allow_module_level=True)

import dgl_b200 as odgl  # noqa: E402
import dgl_b200.nn as onn  # noqa: E402

N, E = 40, 220


def graphs(seed=0, self_loop=True):
    g = torch.Generator().manual_seed(seed)
    u = torch.randint(0, N, (E,), generator=g)
    v = torch.randint(0, N, (E,), generator=g)
    if self_loop:
        u = torch.cat([u, torch.arange(N)])
        v = torch.cat([v, torch.arange(N)])
    return odgl.graph((u, v), num_nodes=N), rdgl.graph((u, v), num_nodes=N)


def build(name, *args, **kw):
    torch.manual_seed(0)
    ref = getattr(rnn, name)(*args, **kw)
    ours = getattr(onn, name)(*args, **kw)
    missing = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return ours.eval(), ref.eval()


def close(a, b, tol=2e-4):
    if isinstance(a, (tuple, list)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            close(x, y, tol)
        return
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a.detach() - b.detach()).abs().max().item()
    assert err <= tol * max(1.0, b.detach().abs().max().item()), err


def run_pair(ours, ref, og, rg, *inputs, extra_o=(), extra_r=(), tol=2e-4, kwargs=None):
    kwargs = kwargs or {}
    xo = [t.clone().requires_grad_(t.is_floating_point()) if torch.is_tensor(t) else t for t in inputs]
    xr = [t.clone().requires_grad_(t.is_floating_point()) if torch.is_tensor(t) else t for t in inputs]
    yo = ours(og, *xo, *extra_o, **kwargs)
    yr = ref(rg, *xr, *extra_r, **kwargs)
    close(yo, yr, tol)
    lo = yo[0] if isinstance(yo, (tuple, list)) else yo
    lr = yr[0] if isinstance(yr, (tuple, list)) else yr
    w = torch.randn_like(lr)
    (lo * w).sum().backward()
    (lr * w).sum().backward()
    for a, b in zip(xo, xr):
        if torch.is_tensor(a) and a.requires_grad and a.grad is not None:
            close(a.grad, b.grad, tol * 5)
    mine = dict(ours.named_parameters())
    for name, p2 in ref.named_parameters():
        if p2.grad is not None:
            assert mine[name].grad is not None, name
            close(mine[name].grad, p2.grad, tol * 5)


X = torch.randn(N, 12, generator=torch.Generator().manual_seed(7))

SIMPLE = [
    ("GraphConv", (12, 8), {"norm": "both"}),
    ("GraphConv", (12, 20), {"norm": "right", "bias": False}),
    ("GraphConv", (12, 8), {"norm": "left"}),
    ("GraphConv", (12, 8), {"norm": "none", "activation": torch.relu}),
    ("SAGEConv", (12, 8, "mean"), {}),
    ("SAGEConv", (12, 8, "gcn"), {}),
    ("SAGEConv", (12, 8, "pool"), {}),
    ("SAGEConv", (12, 8, "lstm"), {}),
    ("GATConv", (12, 6, 3), {}),
    ("GATConv", (12, 6, 3), {"residual": True, "activation": torch.tanh}),
    ("GATv2Conv", (12, 6, 3), {}),
    ("GATv2Conv", (12, 6, 3), {"share_weights": True, "residual": True}),
    ("EGATConv", None, None),
    ("EdgeGATConv", None, None),
    ("DotGatConv", (12, 6, 2), {}),
    ("AGNNConv", (), {}),
    ("TAGConv", (12, 8), {"k": 3}),
    ("SGConv", (12, 8), {"k": 2}),
    ("APPNPConv", (3, 0.1), {}),
    ("GINConv", None, None),
    ("GINEConv", None, None),
    ("ChebConv", (12, 8, 3), {}),
    ("GCN2Conv", (12, 1), {"alpha": 0.3, "lambda_": 0.7}),
    ("GCN2Conv", (12, 2), {"alpha": 0.3, "lambda_": 0.7, "project_initial_features": False}),
    ("GatedGraphConv", (12, 12, 2, 3), {}),
    ("GatedGCNConv", None, None),
    ("GMMConv", None, None),
    ("NNConv", None, None),
    ("CFConv", None, None),
    ("EdgeConv", (12, 8), {}),
    ("EdgeConv", (12, 8), {"batch_norm": True}),
    ("EGNNConv", None, None),
    ("PNAConv", None, None),
    ("DGNConv", None, None),
    ("TWIRLSConv", None, None),
    ("DenseGraphConv", None, None),
    ("DenseSAGEConv", None, None),
    ("DenseChebConv", None, None),
]


@pytest.mark.parametrize("idx", [i for i, c in enumerate(SIMPLE) if c[1] is not None],
                         ids=[f"{c[0]}{i}" for i, c in enumerate(SIMPLE) if c[1] is not None])
def test_conv_layers(idx):
    name, args, kw = SIMPLE[idx]
    og, rg = graphs()
    ours, ref = build(name, *args, **kw)
    extra_o, extra_r, inputs = (), (), (X,)
    if name == "ChebConv":
        extra_o = extra_r = (torch.tensor([2.0] * 1),)
    if name == "GCN2Conv":
        inputs = (X, X * 0.5)
    if name == "GatedGraphConv":
        et = torch.randint(0, 3, (og.num_edges(),), generator=torch.Generator().manual_seed(3))
        extra_o = extra_r = (et,)
    run_pair(ours, ref, og, rg, *inputs, extra_o=extra_o, extra_r=extra_r)


def test_gin_family():
    og, rg = graphs()
    for agg in ("sum", "mean", "max"):
        torch.manual_seed(0)
        ref = rnn.GINConv(torch.nn.Linear(12, 8), agg, init_eps=0.2, learn_eps=True)
        ours = onn.GINConv(torch.nn.Linear(12, 8), agg, init_eps=0.2, learn_eps=True)
        ours.load_state_dict(ref.state_dict())
        run_pair(ours, ref, og, rg, X)
    torch.manual_seed(0)
    ref = rnn.GINEConv(torch.nn.Linear(12, 8), init_eps=0.1, learn_eps=True)
    ours = onn.GINEConv(torch.nn.Linear(12, 8), init_eps=0.1, learn_eps=True)
    ours.load_state_dict(ref.state_dict())
    ef = torch.randn(og.num_edges(), 12)
    run_pair(ours, ref, og, rg, X, ef)


def test_edge_feature_layers():
    og, rg = graphs()
    ne = og.num_edges()
    ef = torch.randn(ne, 5, generator=torch.Generator().manual_seed(1))
    ours, ref = build("EGATConv", 12, 5, 6, 4, 3)
    run_pair(ours, ref, og, rg, X, ef)
    ours, ref = build("EdgeGATConv", 12, 5, 6, 3)
    run_pair(ours, ref, og, rg, X, ef)
    ours, ref = build("GatedGCNConv", 12, 5, 8)
    run_pair(ours, ref, og, rg, X, ef)
    ours, ref = build("GMMConv", 12, 8, 5, 3, "mean")
    run_pair(ours, ref, og, rg, X, ef)
    torch.manual_seed(0)
    fr, fo = torch.nn.Linear(5, 12 * 8), torch.nn.Linear(5, 12 * 8)
    fo.load_state_dict(fr.state_dict())
    ref, ours = rnn.NNConv(12, 8, fr, "sum", residual=True), onn.NNConv(12, 8, fo, "sum", residual=True)
    ours.load_state_dict(ref.state_dict())
    run_pair(ours, ref, og, rg, X, ef)
    ours, ref = build("CFConv", 12, 5, 7, 8)
    run_pair(ours, ref, og, rg, X, ef)
    coord = torch.randn(N, 3, generator=torch.Generator().manual_seed(2))
    ours, ref = build("EGNNConv", 12, 9, 8, 5)
    run_pair(ours, ref, og, rg, X, coord, ef)
    ours, ref = build("PNAConv", 12, 12, ["mean", "max", "sum", "std"], ["identity", "amplification"], 2.5, edge_feat_size=5)
    run_pair(ours, ref, og, rg, X, kwargs={"edge_feat": ef})
    ours, ref = build("DGNConv", 12, 12, ["dir1-av", "dir1-dx", "sum"], ["identity", "amplification"], 2.5)
    ev = torch.randn(N, 2, generator=torch.Generator().manual_seed(5))
    run_pair(ours, ref, og, rg, X, kwargs={"eig_vec": ev})


def test_twirls_and_dense():
    og, rg = graphs()
    ours, ref = build("TWIRLSConv", 12, 5, 16, prop_step=4, dropout=0.0, attn_dropout=0.0)
    run_pair(ours, ref, og, rg, X, tol=1e-3)
    A = torch.zeros(N, N)
    u, v = og.edges()
    A[v, u] = 1.0
    for name, args in (("DenseGraphConv", (12, 8)), ("DenseSAGEConv", (12, 8)), ("DenseChebConv", (12, 8, 3))):
        ours, ref = build(name, *args)
        extra = (2.0,) if name == "DenseChebConv" else ()   # the reference's default calls the removed torch.eig
        close(ours(A, X, *extra), ref(A, X, *extra))


def test_relgraphconv_and_typed_linear():
    og, rg = graphs(self_loop=False)
    et = torch.randint(0, 4, (og.num_edges(),), generator=torch.Generator().manual_seed(3))
    norm = torch.rand(og.num_edges(), 1, generator=torch.Generator().manual_seed(4))
    for reg, nb in ((None, None), ("basis", 2), ("bdd", 2)):
        ours, ref = build("RelGraphConv", 12, 8, 4, regularizer=reg, num_bases=nb, layer_norm=True)
        run_pair(ours, ref, og, rg, X, et, kwargs={"norm": norm})
    xt = torch.randint(0, 5, (N,), generator=torch.Generator().manual_seed(6))
    for reg, nb in ((None, None), ("basis", 3), ("bdd", 4)):
        ours, ref = build("TypedLinear", 12, 8, 5, regularizer=reg, num_bases=nb)
        close(ours(X, xt), ref(X, xt))
        srt, order = torch.sort(xt)
        close(ours(X[order], srt, sorted_by_type=True), ref(X[order], srt, sorted_by_type=True))


def test_hgt_and_hetero_containers():
    og, rg = graphs(self_loop=False)
    nt = torch.randint(0, 3, (N,), generator=torch.Generator().manual_seed(1))
    et = torch.randint(0, 4, (og.num_edges(),), generator=torch.Generator().manual_seed(2))
    ours, ref = build("HGTConv", 12, 4, 3, 3, 4, dropout=0.0, use_norm=True)
    run_pair(ours, ref, og, rg, X, nt, et)
    data = {("a", "r1", "b"): (torch.tensor([0, 1, 2, 3]), torch.tensor([1, 2, 0, 0])),
            ("b", "r2", "a"): (torch.tensor([0, 1, 2]), torch.tensor([3, 0, 1])),
            ("a", "r3", "a"): (torch.tensor([0, 1, 3]), torch.tensor([1, 2, 0]))}
    oh, rh = odgl.heterograph(data), rdgl.heterograph(data)
    feats = {"a": torch.randn(4, 6), "b": torch.randn(3, 6)}
    for agg in ("sum", "mean", "max", "stack"):
        torch.manual_seed(0)
        ref = rnn.HeteroGraphConv({r: rnn.GraphConv(6, 7) for r in ("r1", "r2", "r3")}, aggregate=agg)
        ours = onn.HeteroGraphConv({r: onn.GraphConv(6, 7) for r in ("r1", "r2", "r3")}, aggregate=agg)
        ours.load_state_dict(ref.state_dict())
        yo, yr = ours(oh, feats), ref(rh, feats)
        assert yo.keys() == yr.keys()
        for k in yr:
            close(yo[k], yr[k])
    ours, ref = build("HeteroLinear", {"a": 6, "b": 6}, 5)
    yo, yr = ours(feats), ref(feats)
    for k in yr:
        close(yo[k], yr[k])
    ours, ref = build("HeteroEmbedding", {"a": 4, "b": 3}, 5)
    ids = {"a": torch.tensor([0, 3]), "b": torch.tensor([1])}
    yo, yr = ours(ids), ref(ids)
    for k in yr:
        close(yo[k], yr[k])


def test_readout_and_utils():
    og, rg = graphs()
    ob, rb = odgl.batch([og, og]), rdgl.batch([rg, rg])
    xx = torch.cat([X, X * 0.5])
    for name, args in (("SumPooling", ()), ("AvgPooling", ()), ("MaxPooling", ()), ("SortPooling", (3,)),
                       ("WeightAndSum", (12,)), ("Set2Set", (12, 2, 1)), ("SetTransformerEncoder", (12, 2, 4, 8)),
                       ("SetTransformerDecoder", (12, 2, 4, 8, 1, 2))):
        ours, ref = build(name, *args)
        close(ours(ob, xx), ref(rb, xx), 5e-4)
    torch.manual_seed(0)
    gr, go = torch.nn.Linear(12, 1), torch.nn.Linear(12, 1)
    go.load_state_dict(gr.state_dict())
    close(onn.GlobalAttentionPooling(go)(ob, xx), rnn.GlobalAttentionPooling(gr)(rb, xx))
    lp_o, lp_r = onn.LabelPropagation(3, 0.8), rnn.LabelPropagation(3, 0.8)
    lab = torch.randint(0, 4, (N,), generator=torch.Generator().manual_seed(9))
    mask = torch.rand(N, generator=torch.Generator().manual_seed(10)) < 0.5
    close(lp_o(og, lab, mask), lp_r(rg, lab, mask))
    for mode in ("cat", "max", "lstm"):
        ours, ref = build("JumpingKnowledge", mode, 12, 2)
        close(ours([X, X * 2]), ref([X, X * 2]))
    ours, ref = build("WeightBasis", (3, 4), 2, 5)
    close(ours(), ref())
    for op in ("dot", "cos", "ele", "cat"):
        ours, ref = build("EdgePredictor", op, 12, 3, bias=True)
        close(ours(X, X.flip(0)), ref(X, X.flip(0)))
    rels = torch.randint(0, 3, (N,), generator=torch.Generator().manual_seed(11))
    ours, ref = build("TransE", 3, 12)
    close(ours(X, X.flip(0), rels), ref(X, X.flip(0), rels))
    ours, ref = build("TransR", 3, 5, 12)
    close(ours(X, X.flip(0), rels), ref(X, X.flip(0), rels))


def test_graph_transformer_blocks():
    import dgl.nn.pytorch.gt as rgt
    import dgl_b200.nn.gt as ogt
    g = torch.Generator().manual_seed(0)
    B, n, d, H = 2, 6, 16, 4
    x = torch.randn(B, n, d, generator=g)
    bias = torch.randn(B, n, n, H, generator=g)
    mask = torch.zeros(B, n, n, dtype=torch.bool)
    mask[1, :, 4:] = True

    def pair(name, *a, **kw):
        torch.manual_seed(0)
        ref = getattr(rgt, name)(*a, **kw)
        ours = getattr(ogt, name)(*a, **kw)
        ours.load_state_dict(ref.state_dict(), strict=True)
        return ours.eval(), ref.eval()

    for btype in ("add", "mul"):
        ours, ref = pair("BiasedMHA", d, H, attn_bias_type=btype, attn_drop=0.0)
        close(ours(x, bias, mask), ref(x, bias, mask))
    for norm_first in (False, True):
        ours, ref = pair("GraphormerLayer", d, 32, H, norm_first=norm_first, dropout=0.0, attn_dropout=0.0)
        close(ours(x, bias, mask), ref(x, bias, mask))
    e = torch.randn(B, n, n, 8, generator=g)
    ours, ref = pair("EGTLayer", d, 8, H, 1)
    close(ours(x, e), ref(x, e))
    fmask = torch.zeros(B, n, n)
    fmask[1, :, 5:] = -1e9
    close(ours(x, e, fmask), ref(x, e, fmask))
    ours, ref = pair("EGTLayer", d, 8, H, 0, edge_update=False)
    close(ours(x, e), ref(x, e))
    deg = torch.randint(0, 9, (2, B, n), generator=g)
    ours, ref = pair("DegreeEncoder", 5, d, "both")
    close(ours(deg), ref(deg))
    ours, ref = pair("DegreeEncoder", 5, d, "in")
    close(ours(deg[0]), ref(deg[0]))
    dist = torch.randint(-1, 6, (B, n, n), generator=g)
    ours, ref = pair("SpatialEncoder", 4, H)
    close(ours(dist), ref(dist))
    path = torch.randn(B, n, n, 3, 5, generator=g)
    ours, ref = pair("PathEncoder", 3, 5, H)
    close(ours(dist, path), ref(dist, path))
    coord = torch.randn(B, n, 3, generator=g)
    ntype = torch.randint(0, 3, (B, n), generator=g)
    ours, ref = pair("SpatialEncoder3d", 4, H, 3)
    close(ours(coord), ref(coord))
    close(ours(coord, ntype), ref(coord, ntype))
    ev = torch.randn(n, 5, generator=g)
    ea = torch.rand(n, 5, generator=g)
    ev[:, 4] = float("nan")
    ea[:, 4] = float("nan")
    for model, kw in (("Transformer", {"n_head": 2}), ("DeepSet", {"num_post_layer": 2}), ("DeepSet", {"batch_norm": True})):
        ours, ref = pair("LapPosEncoder", model, 3 if model == "DeepSet" else 1, 5, 8, **kw)
        close(ours(ea.clone(), ev.clone()), ref(ea.clone(), ev.clone()), 1e-3)


def test_homophily_measures_match_reference():
    """node / edge / LINKX / adjusted homophily vs the installed reference on a random labelled graph (and the docstring
    example of the adjusted measure, which symmetrises the graph first)."""
    torch.manual_seed(3)
    u, v = torch.randint(0, 150, (2500,)), torch.randint(0, 150, (2500,))
    y = torch.randint(0, 6, (150,))
    go, gr = odgl.graph((u, v), num_nodes=150), rdgl.graph((u, v), num_nodes=150)
    for name in ("node_homophily", "edge_homophily", "linkx_homophily", "adjusted_homophily"):
        assert abs(getattr(odgl, name)(go, y) - getattr(rdgl, name)(gr, y)) < 1e-6, name
    tiny = odgl.graph(([1, 2, 0, 4], [0, 1, 2, 3]))
    assert abs(odgl.adjusted_homophily(tiny, torch.tensor([0, 0, 0, 0, 1])) + 0.1428571492433548) < 1e-6
