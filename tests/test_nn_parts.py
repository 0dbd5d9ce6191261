"""Conv building blocks (reference nn/pytorch/conv/{twirlsconv,grouprevres,cugraph_base}.py) against dense math."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200.nn.pytorch.conv import (AX, MLP, Attention, CuGraphBaseConv, D_power_bias_X, D_power_X,
                                      InvertibleCheckpoint, Propagate, PropagateNoPrecond, normalized_AX)
from dgl_b200.nn.pytorch.factory import pairwise_squared_distance
from dgl_b200.nn.pytorch.gt import gaussian


def _graph():
    g = dgl.rand_graph(30, 150)
    g = dgl.to_simple(dgl.add_reverse_edges(g))
    g.edata["w"] = torch.rand(g.num_edges()) + 0.1
    u, v = g.edges()
    A = torch.zeros(30, 30).index_put((v.long(), u.long()), g.edata["w"])
    g.ndata["deg"] = A.sum(1) + 0.5
    return g, A


def test_propagation_steps_match_dense_equations():
    g, A = _graph()
    X, Y = torch.randn(30, 5), torch.randn(30, 5)
    d = g.ndata["deg"]
    assert torch.allclose(AX(g, X), A @ X, atol=1e-5)
    assert torch.allclose(D_power_X(g, X, -0.5), X / d.sqrt().unsqueeze(1), atol=1e-6)
    assert torch.allclose(D_power_bias_X(g, X, -1, 0.3, 0.7), X / (0.3 * d + 0.7).unsqueeze(1), atol=1e-6)
    Dm = torch.diag(d.pow(-0.5))
    assert torch.allclose(normalized_AX(g, X), Dm @ A @ Dm @ X, atol=1e-5)
    alp, lam = 0.4, 0.6
    Dt = lam * d + (1 - lam)
    ref = (1 - alp) * Y + alp * lam * (torch.diag(Dt.pow(-0.5)) @ A @ torch.diag(Dt.pow(-0.5)) @ Y) + alp * X / Dt.unsqueeze(1)
    assert torch.allclose(Propagate()(g, Y, X, alp, lam), ref, atol=1e-5)
    ref2 = (1 - alp * lam - alp) * Y + alp * lam * (Dm @ A @ Dm @ Y) + alp * X
    assert torch.allclose(PropagateNoPrecond()(g, Y, X, alp, lam), ref2, atol=1e-5)


def test_irls_attention_reweights_and_refreshes_degrees():
    g, _ = _graph()
    Y = torch.randn(30, 4)
    att = Attention(tau=0.3, T=2.0, p=1.0)
    att(g, Y)
    u, v = g.edges()
    dist = (Y[u.long()] - Y[v.long()]).norm(dim=1)                        # ||.||^{2-p} with p = 1
    w = dist.clamp(min=0.3)
    expect = torch.where(w > 2.0, torch.zeros_like(w), 1 / w)
    assert torch.allclose(g.edata["w"], expect + 1e-9, atol=1e-4)
    deg = torch.zeros(30).index_add_(0, v.long(), g.edata["w"])
    assert torch.allclose(g.ndata["deg"], deg, atol=1e-5)
    etas = torch.tensor([1.0, 0.0, 0.0, 0.0])
    Attention(0.01, -1, 2.0)(g, Y, etas)                                    # p = 2 -> all weights 1/max(tau, 1)
    assert torch.allclose(g.edata["w"], torch.ones(g.num_edges()), atol=1e-6)
    drop = Attention(0.3, -1, 1.0, attn_dropout=0.5).train()
    drop(g, Y)
    assert (g.edata["w"] == 0).any()


def test_mlp_shapes_and_norm_counts():
    m = MLP(6, 8, 3, num_layers=3, dropout=0.0, norm="batch", init_activate=False)
    assert len(m.layers) == 3 and len(m.norms) == 2 and m(torch.randn(10, 6)).shape == (10, 3)
    m = MLP(8, 8, 3, num_layers=2, dropout=0.1, norm="layer", init_activate=True)
    assert len(m.norms) == 2 and m(torch.randn(4, 8)).shape == (4, 3)
    assert MLP(5, 7, 2, 1, 0.0, "none", False).norms is None
    ident = MLP(5, 5, 5, 0, 0.0, "none", False)
    x = torch.randn(3, 5)
    assert torch.equal(ident(x), x)


def test_invertible_checkpoint_gradients_equal_plain_autograd():
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 4)

    def f(x, c):
        return x + torch.tanh(lin(c))                 # invertible in x given c

    def f_inv(y, c):
        return y - torch.tanh(lin(c))

    x0 = torch.randn(6, 4, requires_grad=True)
    c = torch.randn(6, 4, requires_grad=True)
    w = torch.randn(6, 4)
    ref = (f(x0 * 1.0, c) * w).sum()
    gx, gc, gw, gb = torch.autograd.grad(ref, [x0, c, lin.weight, lin.bias])
    x1 = x0.detach().clone().requires_grad_(True)
    xin = x1 * 1.0                                      # non-leaf: its storage may be freed by the checkpoint
    out = InvertibleCheckpoint.apply(f, f_inv, 2, xin, c, lin.weight, lin.bias)
    assert xin.untyped_storage().size() == 0            # activation memory released after forward
    (out * w).sum().backward()
    assert torch.allclose(x1.grad, gx, atol=1e-6) and torch.allclose(c.grad, gc, atol=1e-6)
    assert torch.allclose(lin.weight.grad, gw, atol=1e-6) and torch.allclose(lin.bias.grad, gb, atol=1e-6)


def test_cugraph_base_pad_offsets_and_small_functions():
    base = CuGraphBaseConv()
    off = torch.tensor([0, 2, 5])
    assert base.pad_offsets(off, 6).tolist() == [0, 2, 5, 5, 5, 5]
    assert base.pad_offsets(torch.tensor([0, 1]), 3).tolist() == [0, 1, 1]       # cached buffer reused
    assert base.pad_offsets(off, 9).numel() == 9
    with pytest.raises(NotImplementedError):
        base.forward()
    x = torch.randn(2, 5, 3)
    assert torch.allclose(pairwise_squared_distance(x), torch.cdist(x, x) ** 2, atol=1e-4)
    g = gaussian(torch.tensor([0.0]), torch.tensor([0.0]), torch.tensor([1.0]))
    assert abs(float(g) - 0.39894) < 1e-3


def test_gat_score_folding_equals_direct_form():
    """``a . (W h) == (W^T a) . h``: the folded logit halves of GATConv equal the direct broadcast-multiply-reduce form,
    for the shared projection and for separate source / destination projections."""
    import torch
    import dgl_b200.nn as dglnn
    torch.manual_seed(0)
    for in_feats in (12, (12, 7)):
        conv = dglnn.GATConv(in_feats, 5, 3)
        n_src, n_dst = 20, 9
        h_src = torch.randn(n_src, 12)
        h_dst = h_src[:n_dst] if isinstance(in_feats, int) else torch.randn(n_dst, 7)
        if isinstance(in_feats, int):
            z_src = conv.fc(h_src).view(n_src, 3, 5)
            z_dst = z_src[:n_dst]
        else:
            z_src, z_dst = conv.fc_src(h_src).view(n_src, 3, 5), conv.fc_dst(h_dst).view(n_dst, 3, 5)
        a, b = conv._scores(h_src, h_dst, z_src, z_dst)
        assert torch.allclose(a, (z_src * conv.attn_l).sum(-1), atol=1e-5)
        assert torch.allclose(b, (z_dst * conv.attn_r).sum(-1), atol=1e-5)
