"""dgl.sparse API + nn layer behaviour (CPU)."""
import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.sparse as dglsp
import dgl_b200.nn as dglnn


def _rand_sp(m=6, n=5, nnz=12, d=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    key = torch.randperm(m * n, generator=g)[:nnz]
    r, c = key // n, key % n
    val = torch.randn((nnz,) if d is None else (nnz, d), generator=g)
    return dglsp.spmatrix(torch.stack([r, c]), val, (m, n))


def test_construct_and_formats():
    A = _rand_sp()
    D = A.to_dense()
    ip, idx, e = A.csr()
    B = dglsp.from_csr(ip, idx, A.val[e.long()] if e is not None else A.val, A.shape)
    assert torch.allclose(B.to_dense(), D)
    ip, idx, e = A.csc()
    C = dglsp.from_csc(ip, idx, A.val[e.long()] if e is not None else A.val, A.shape)
    assert torch.allclose(C.to_dense(), D)
    assert torch.allclose(A.t().to_dense(), D.t())
    assert torch.allclose(dglsp.identity((3, 3)).to_dense(), torch.eye(3))
    assert A.nnz == 12 and A.shape == (6, 5)


def test_spmm_sddmm_grad():
    A = _rand_sp()
    A.val.requires_grad_()
    X = torch.randn(5, 4, requires_grad=True)
    Y = dglsp.spmm(A, X)
    ref = A.to_dense() @ X
    assert torch.allclose(Y, ref, atol=1e-5)
    g1 = torch.autograd.grad(Y.sum(), [A.val, X], retain_graph=True)
    g2 = torch.autograd.grad(ref.sum(), [A.val, X])
    assert all(torch.allclose(a, b, atol=1e-5) for a, b in zip(g1, g2))
    X1, X2 = torch.randn(6, 3), torch.randn(3, 5)
    S = dglsp.sddmm(A, X1, X2)
    mask = (A.to_dense() != 0).float()
    assert torch.allclose(S.to_dense(), A.to_dense() * (X1 @ X2) * mask, atol=1e-5)
    assert torch.allclose((A @ X), ref, atol=1e-5)


def test_spspmm_softmax_reduce_elementwise():
    A, B = _rand_sp(6, 5, 12, seed=1), _rand_sp(5, 4, 9, seed=2)
    C = dglsp.spspmm(A, B)
    assert torch.allclose(C.to_dense(), A.to_dense() @ B.to_dense(), atol=1e-5)
    S = dglsp.softmax(A)
    d = A.to_dense()
    m = d != 0
    ref = torch.where(m, d, torch.full_like(d, float("-inf"))).softmax(1)
    ref = torch.where(m, ref, torch.zeros_like(ref))
    assert torch.allclose(S.to_dense(), ref, atol=1e-5)
    assert torch.allclose(A.sum(1), d.sum(1), atol=1e-5)
    assert torch.allclose(A.sum(0), d.sum(0), atol=1e-5)
    assert torch.allclose((A + A).to_dense(), 2 * d, atol=1e-6)
    assert torch.allclose((A * 3).to_dense(), 3 * d)
    assert torch.allclose((A / 2).to_dense(), d / 2)
    assert torch.allclose(dglsp.sp_mul_v(A, torch.arange(5.0)).to_dense(), d * torch.arange(5.0))
    Dg = dglsp.diag(torch.tensor([1.0, 2, 4]))
    assert torch.allclose(Dg.inv().to_dense(), torch.diag(torch.tensor([1.0, 0.5, 0.25])))
    Bv = _rand_sp(6, 5, 12, d=3)
    X = torch.randn(5, 2, 3)
    out = dglsp.bspmm(Bv, X)
    ref = torch.stack([Bv.to_dense()[..., b] @ X[..., b] for b in range(3)], -1)
    assert torch.allclose(out, ref, atol=1e-5)


def test_layers_grad_and_equivalence():
    torch.manual_seed(0)
    g = dgl.add_self_loop(dgl.rand_graph(30, 150))
    x = torch.randn(30, 6)
    # GraphConv == dense normalised adjacency
    conv = dglnn.GraphConv(6, 4, norm="both", bias=False)
    u, v = g.edges()
    A = torch.zeros(30, 30).index_put_((v, u), torch.ones(len(u)), accumulate=True)
    dout = g.out_degrees().float().clamp(min=1).pow(-0.5)
    din = g.in_degrees().float().clamp(min=1).pow(-0.5)
    ref = din.unsqueeze(1) * (A @ (dout.unsqueeze(1) * x)) @ conv.weight
    assert torch.allclose(conv(g, x), ref, atol=1e-4)
    # SAGE mean
    sage = dglnn.SAGEConv(6, 4, "mean")
    deg = g.in_degrees().float().clamp(min=1).unsqueeze(1)
    ref = sage.fc_self(x) + sage.fc_neigh((A @ x) / deg)
    assert torch.allclose(sage(g, x), ref, atol=1e-4)
    # GAT gradient flows to all parameters
    gat = dglnn.GATConv(6, 4, 2)
    gat(g, x).sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in gat.parameters())
    # RelGraphConv low-mem vs dense per-relation
    et = torch.randint(0, 3, (g.num_edges(),))
    rel = dglnn.RelGraphConv(6, 4, 3, self_loop=False, bias=False)
    W = rel.linear_r.get_weight()
    ref = torch.zeros(30, 4)
    ref.index_add_(0, v, torch.einsum("ek,ekm->em", x[u], W[et]))
    assert torch.allclose(rel(g, x, et), ref, atol=1e-4)
    # sorted path gives same result
    order = torch.argsort(et)
    g2 = dgl.graph((u[order], v[order]), num_nodes=30)
    assert torch.allclose(rel(g2, x, et[order], presorted=True), ref, atol=1e-4)
    with pytest.raises(dgl.DGLError):
        dglnn.GraphConv(6, 4)(dgl.graph(([0], [1]), num_nodes=3), torch.randn(3, 6))
