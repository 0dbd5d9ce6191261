"""Behaviour of the explanation modules (parity targets: /root/reference/tests/python/pytorch/nn/test_nn.py gnnexplainer /
pgexplainer / subgraphx cases, plus value-level checks the reference suite does not make)."""
import math
import os
import sys

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import dgl_b200 as dgl
from dgl_b200.nn import explain as X
from dgl_b200.nn import GraphConv, HeteroGraphConv, AvgPooling

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def ring_with_chords(n=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    u = torch.arange(n)
    v = (u + 1) % n
    cu = torch.randint(0, n, (n // 2,), generator=g)
    cv = (cu + 3) % n
    src, dst = torch.cat([u, v, cu, cv]), torch.cat([v, u, cv, cu])
    return dgl.graph((src, dst), num_nodes=n)


class NodeModel(nn.Module):
    def __init__(self, d, c):
        super().__init__()
        self.conv = GraphConv(d, c, allow_zero_in_degree=True)

    def forward(self, graph, feat, eweight=None):
        return self.conv(graph, feat, edge_weight=eweight)


class GraphModel(nn.Module):
    def __init__(self, d, c):
        super().__init__()
        self.conv = GraphConv(d, c, allow_zero_in_degree=True)
        self.pool = AvgPooling()

    def forward(self, graph, feat, eweight=None):
        return self.pool(graph, self.conv(graph, feat, edge_weight=eweight))


def test_gnnexplainer_node_and_graph():
    torch.manual_seed(0)
    g = ring_with_chords()
    x = torch.randn(g.num_nodes(), 5)
    ex = X.GNNExplainer(NodeModel(5, 3), num_hops=1, num_epochs=20, lr=0.05, log=False)
    centre, sg, fmask, emask = ex.explain_node(4, g, x)
    assert sg.ndata[dgl.NID][centre].item() == 4
    assert fmask.shape == (5,) and emask.shape == (sg.num_edges(),)
    assert ((fmask > 0) & (fmask < 1)).all() and ((emask > 0) & (emask < 1)).all()
    ex = X.GNNExplainer(GraphModel(5, 3), num_hops=1, num_epochs=20, log=False)
    fmask, emask = ex.explain_graph(g, x)
    assert fmask.shape == (5,) and emask.shape == (g.num_edges(),)


def test_gnnexplainer_objective_decreases():
    torch.manual_seed(1)
    g = ring_with_chords()
    x = torch.randn(g.num_nodes(), 6)
    model = GraphModel(6, 4)
    ex = X.GNNExplainer(model, num_hops=1, num_epochs=0, log=False)
    masks = X._TypedMasks(g, {"_N": x})
    nt, rel = g.ntypes[0], g.canonical_etypes[0]
    target = model(g, x).argmax(-1)

    def total():
        logp = model(g, x * masks.feat[nt].sigmoid(), eweight=masks.edge[rel].sigmoid()).log_softmax(-1)
        return (-logp[0, target[0]] + masks.penalty(masks.feat.values(), masks.edge.values(), *ex._coeffs())).item()

    before = total()
    ex.num_epochs = 60
    ex.lr = 0.1
    ex._optimise(masks, lambda fg, eg: model(g, x * fg[nt], eweight=eg[rel]), lambda logp: -logp[0, target[0]])
    assert total() < before


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dgl")), reason="reference install not present")
def test_gnnexplainer_regulariser_matches_reference():
    sys.path.insert(0, REF)
    try:
        from dgl.nn.pytorch.explain import GNNExplainer as RefEx
    finally:
        sys.path.remove(REF)
    torch.manual_seed(2)
    fm, em = torch.randn(1, 7), torch.randn(31)
    ours = X.GNNExplainer(None, 2, alpha1=0.01, alpha2=0.7, beta1=0.9, beta2=0.2)._loss_regularize(torch.tensor(0.3), fm, em)
    ref = RefEx(nn.Identity(), 2, alpha1=0.01, alpha2=0.7, beta1=0.9, beta2=0.2)._loss_regularize(torch.tensor(0.3), fm, em)
    assert torch.allclose(ours, ref, atol=1e-6)


def hetero_toy():
    g = dgl.heterograph({("user", "follows", "user"): (torch.tensor([0, 1, 2, 3]), torch.tensor([1, 2, 3, 0])),
                         ("user", "plays", "game"): (torch.tensor([0, 1, 2, 3, 0]), torch.tensor([0, 1, 0, 1, 1])),
                         ("game", "played-by", "user"): (torch.tensor([0, 1, 0, 1, 1]), torch.tensor([0, 1, 2, 3, 0]))})
    feat = {"user": torch.randn(4, 5), "game": torch.randn(2, 5)}
    return g, feat


class HeteroModel(nn.Module):
    def __init__(self, d, c, rels, graph_level):
        super().__init__()
        self.conv = HeteroGraphConv({r: GraphConv(d, c, allow_zero_in_degree=True) for r in rels})
        self.graph_level = graph_level

    def forward(self, graph, feat, eweight=None, embed=False, edge_weight=None):
        w = eweight if eweight is not None else edge_weight
        kw = {} if w is None else {"mod_kwargs": {c[1]: {"edge_weight": w[c]} for c in graph.canonical_etypes}}
        h = self.conv(graph, feat, **kw)
        if embed or not self.graph_level:
            return h
        with graph.local_scope():
            graph.ndata["h"] = h
            return sum(dgl.mean_nodes(graph, "h", ntype=nt) for nt in h)


def test_hetero_gnnexplainer():
    torch.manual_seed(3)
    g, feat = hetero_toy()
    model = HeteroModel(5, 3, g.etypes, graph_level=True)
    fm, em = X.HeteroGNNExplainer(model, num_hops=1, num_epochs=5, log=False).explain_graph(g, feat)
    assert set(fm) == set(g.ntypes) and set(em) == set(g.canonical_etypes)
    for c in g.canonical_etypes:
        assert em[c].shape == (g.num_edges(c),)
    model = HeteroModel(5, 3, g.etypes, graph_level=False)
    centre, sg, fm, em = X.HeteroGNNExplainer(model, num_hops=1, num_epochs=5, log=False).explain_node("user", 0, g, feat)
    assert sg.nodes["user"].data[dgl.NID][centre].item() == 0
    assert all(em[c].shape == (sg.num_edges(c),) for c in sg.canonical_etypes)


class PGModel(nn.Module):
    def __init__(self, d, c, graph_level):
        super().__init__()
        self.conv = GraphConv(d, c, allow_zero_in_degree=True)
        self.pool = AvgPooling()
        self.graph_level = graph_level

    def forward(self, g, h, embed=False, edge_weight=None):
        h = self.conv(g, h, edge_weight=edge_weight)
        if embed or not self.graph_level:
            return h
        return self.pool(g, h)


def test_reverse_values_lookup():
    u = torch.tensor([0, 1, 2, 2, 3])
    v = torch.tensor([1, 0, 3, 0, 2])
    vals = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0])
    # (0,1)<->(1,0), (2,3)<->(3,2); (2,0) has no reverse and keeps its own value
    assert X._reverse_values(u, v, 4, vals).tolist() == [2.0, 1.0, 5.0, 4.0, 3.0]


def test_pgexplainer_graph_level_training_and_symmetry():
    torch.manual_seed(4)
    g = ring_with_chords()
    x = torch.randn(g.num_nodes(), 5)
    ex = X.PGExplainer(PGModel(5, 3, True), 3)
    opt = torch.optim.Adam(ex.parameters(), lr=0.01)
    assert len(list(ex.parameters())) == 4
    for t in (5.0, 2.0):
        loss = ex.train_step(g, x, t)
        assert loss.dim() == 0 and torch.isfinite(loss)
        opt.zero_grad()
        loss.backward()
        assert all(p.grad is not None for p in ex.parameters())
        opt.step()
    probs, mask = ex.explain_graph(g, x)
    assert probs.shape == (1, 3) and mask.shape == (g.num_edges(),)
    u, v = g.edges()
    lut = {(a, b): m for a, b, m in zip(u.tolist(), v.tolist(), mask.tolist())}
    assert all(abs(lut[(a, b)] - lut[(b, a)]) < 1e-6 for (a, b) in lut if (b, a) in lut)
    assert ex.edge_mask is None
    with pytest.raises(AssertionError):
        ex.explain_node(0, g, x)


def test_pgexplainer_node_level():
    torch.manual_seed(5)
    g = ring_with_chords()
    x = torch.randn(g.num_nodes(), 5)
    ex = X.PGExplainer(PGModel(5, 3, False), 3, num_hops=1, explain_graph=False)
    loss = ex.train_step_node(torch.tensor([0, 5]), g, x, 3.0)
    loss.backward()
    assert torch.isfinite(loss)
    probs, mask, bg, centres = ex.explain_node([0, 5], g, x)
    assert bg.batch_size == 2 and centres.numel() == 2
    assert bg.ndata[dgl.NID][centres].tolist() == [0, 5]
    assert mask.shape == (bg.num_edges(),) and probs.shape == (bg.num_nodes(), 3)


def test_hetero_pgexplainer():
    torch.manual_seed(6)
    g, feat = hetero_toy()
    ex = X.HeteroPGExplainer(HeteroModel(5, 3, g.etypes, graph_level=True), 3)
    loss = ex.train_step(g, feat, 2.0)
    loss.backward()
    probs, masks = ex.explain_graph(g, feat)
    assert probs.shape == (1, 3)
    assert {c: m.shape[0] for c, m in masks.items()} == {c: g.num_edges(c) for c in g.canonical_etypes}
    ex = X.HeteroPGExplainer(HeteroModel(5, 3, g.etypes, graph_level=False), 3, num_hops=1, explain_graph=False)
    loss = ex.train_step_node({"user": [0, 1]}, g, feat, 2.0)
    loss.backward()
    probs, masks, bg, centres = ex.explain_node({"user": [0, 1]}, g, feat)
    assert centres["user"].numel() == 2 and bg.nodes["user"].data[dgl.NID][centres["user"]].tolist() == [0, 1]
    assert all(masks[c].shape[0] == bg.num_edges(c) for c in bg.canonical_etypes)


def test_component_labels():
    u = torch.tensor([0, 1, 3, 5, 6])
    v = torch.tensor([1, 2, 4, 6, 7])
    lab = X._component_labels(9, u, v)
    assert lab.tolist() == [0, 0, 0, 3, 3, 5, 5, 5, 8]
    assert X._largest_component(9, u, v).tolist() == [0, 1, 2]
    assert X._largest_component(9, u, v, without=1).tolist() in ([5, 6, 7],)


class MotifModel(nn.Module):
    """Graph logit for class 1 grows with the (neighbour-averaged) signal carried by the planted nodes."""

    def forward(self, g, h):
        with g.local_scope():
            g.ndata["h"] = h
            g.update_all(dgl.function.copy_u("h", "m"), dgl.function.mean("m", "a"))
            s = dgl.sum_nodes(g, "a")
        return torch.cat([torch.zeros_like(s), 4.0 * s], dim=1)


def test_subgraphx_finds_planted_nodes_and_batched_value_matches():
    torch.manual_seed(7)
    g = ring_with_chords(10, seed=1)
    x = torch.zeros(10, 1)
    x[[2, 3, 4]] = 1.0
    ex = X.SubgraphX(MotifModel(), num_hops=2, num_rollouts=12, node_min=3, shapley_steps=30, num_child=6)
    nodes = ex.explain_graph(g, x, target_class=1)
    assert len(nodes) <= 3
    assert len(set(nodes.tolist()) & {2, 3, 4}) >= 2
    # the batched coalition evaluation computes the same estimator
    torch.manual_seed(11)
    a = ex.shapley(torch.tensor([2, 3, 4]))
    ex.coalition_batch = 8
    torch.manual_seed(11)
    b = ex.shapley(torch.tensor([2, 3, 4]))
    assert math.isclose(a, b, rel_tol=1e-5, abs_tol=1e-6)
    assert a > ex.shapley(torch.tensor([7, 8, 9])) - 1e-9


def test_hetero_subgraphx_runs():
    torch.manual_seed(8)
    g, feat = hetero_toy()
    model = HeteroModel(5, 3, g.etypes, graph_level=True)
    out = X.HeteroSubgraphX(model, num_hops=1, num_rollouts=3, node_min=2, shapley_steps=5).explain_graph(g, feat, 1)
    assert set(out) == set(g.ntypes) and sum(v.numel() for v in out.values()) <= 2


# ------------------------------------------------------------------------------------------- skip-gram embedding models
def test_deepwalk_window_pairs_and_training():
    from dgl_b200.nn.network_emb import DeepWalk, _window_offsets
    ctx, centre = _window_offsets(5, 2)
    pairs = set(zip(ctx.tolist(), centre.tolist()))
    assert pairs == {(i, j) for i in range(5) for j in range(5) if i != j and abs(i - j) <= 2}
    torch.manual_seed(0)
    g = ring_with_chords(16)
    model = DeepWalk(g, emb_dim=8, walk_length=6, window_size=2, negative_size=2, sparse=False)
    assert model.context_embed.weight.abs().sum() == 0 and model.node_embed.weight.abs().max() <= 1 / 8
    walks = model.sample(torch.arange(16))
    assert walks.shape == (16, 6)
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    first = None
    for _ in range(30):
        loss = model(walks)
        first = first if first is not None else loss.item()
        opt.zero_grad()
        loss.backward()
        opt.step()
    assert loss.item() < first
    slow = DeepWalk(g, emb_dim=8, walk_length=6, window_size=2, negative_size=2, fast_neg=False, sparse=False)
    assert torch.isfinite(slow(walks))
    with pytest.raises(AssertionError):
        DeepWalk(g, walk_length=3, window_size=5)


def test_metapath2vec_sampling_ids_and_loss():
    from dgl_b200.nn.network_emb import MetaPath2Vec
    torch.manual_seed(1)
    g = dgl.heterograph({("user", "uc", "company"): (torch.tensor([0, 1, 2, 3]), torch.tensor([0, 0, 1, 1])),
                         ("company", "cu", "user"): (torch.tensor([0, 0, 1, 1]), torch.tensor([0, 1, 2, 3]))})
    model = MetaPath2Vec(g, ["uc", "cu"], window_size=1, emb_dim=8, negative_size=3, sparse=False)
    assert model.node_metapath == ["user", "company", "user"]
    n_user, n_comp = 4, 2
    assert model.node_embed.num_embeddings == n_user + n_comp
    l2g = model.local_to_global_nid
    assert set(l2g) == {"user", "company"} and len(set(l2g["user"].tolist()) | set(l2g["company"].tolist())) == 6
    assert abs(model.neg_prob.sum().item() - 1.0) < 1e-5
    pos_u, pos_v, neg_v = model.sample(torch.arange(4))
    assert pos_u.shape == pos_v.shape and neg_v.shape == (pos_u.numel(), 3)
    # window 1 on a user-company-user walk: every pair joins a user and a company
    users, comps = set(l2g["user"].tolist()), set(l2g["company"].tolist())
    for a, b in zip(pos_u.tolist(), pos_v.tolist()):
        assert (a in users) != (b in users) and (a in comps) != (b in comps)
    assert pos_u.numel() == 4 * 4          # positions (0,1),(1,0),(1,2),(2,1) per walk
    loss = model(pos_u, pos_v, neg_v)
    loss.backward()
    assert torch.isfinite(loss) and model.node_embed.weight.grad is not None
