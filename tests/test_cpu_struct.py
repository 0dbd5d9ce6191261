"""Native host SpGEMM / sparse add / masked row-dot (csrc/cpu/structure.cc) against dense PyTorch references.

Modelled on the reference's tests/python/common/ops/test_ops.py::test_csrmm / test_csrsum / test_csrmask."""
import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.sparse as dglsp
from dgl_b200.ops import cpu_struct


def _rand(m, n, nnz, dtype=torch.float32, idtype=torch.int64, requires_grad=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    key = torch.unique(torch.randint(0, m * n, (nnz,), generator=g))
    key = key[torch.randperm(key.numel(), generator=g)]
    r, c = (key // n).to(idtype), (key % n).to(idtype)
    v = torch.randn(key.numel(), generator=g).to(dtype).requires_grad_(requires_grad)
    return dglsp.from_coo(r, c, v, (m, n)), v


@pytest.mark.parametrize("idtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_spspmm_matches_dense_and_backward(idtype, dtype):
    A, va = _rand(37, 29, 200, dtype, idtype, True, 1)
    B, vb = _rand(29, 41, 260, dtype, idtype, True, 2)
    assert cpu_struct.usable(A, B)
    C = dglsp.spspmm(A, B)
    ref = A.to_dense() @ B.to_dense()
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert torch.allclose(C.to_dense(), ref, atol=tol)
    # output rows are column-sorted and free of duplicates
    ip, idx, _ = C.csr()
    for r in range(C.shape[0]):
        row = idx[ip[r]:ip[r + 1]]
        assert torch.all(row[1:] > row[:-1])
    w = torch.randn(37, 41, dtype=dtype)
    (C.to_dense() * w).sum().backward()
    ga, gb = va.grad.clone(), vb.grad.clone()
    va.grad = vb.grad = None
    (ref * w).sum().backward()
    assert torch.allclose(ga, va.grad, atol=tol * 10)
    assert torch.allclose(gb, vb.grad, atol=tol * 10)


def test_spspmm_low_precision_and_empty_rows():
    A, _ = _rand(20, 30, 40, torch.bfloat16, seed=3)
    B, _ = _rand(30, 10, 25, torch.bfloat16, seed=4)
    C = dglsp.spspmm(A, B)
    assert C.val.dtype == torch.bfloat16
    assert torch.allclose(C.to_dense().float(), A.to_dense().float() @ B.to_dense().float(), atol=5e-2, rtol=5e-2)
    E = dglsp.from_coo(torch.tensor([], dtype=torch.int64), torch.tensor([], dtype=torch.int64), torch.zeros(0), (20, 30))
    Z = dglsp.spspmm(E, B.to(dtype=torch.float32))
    assert Z.nnz == 0 and Z.shape == (20, 10)


def test_repeated_calls_with_changing_widths_share_thread_scratch():
    # the per-thread marker array is reused across calls; results must not depend on call history
    for n, seed in [(500, 5), (50, 6), (300, 7), (50, 6)]:
        A, _ = _rand(64, 64, 400, seed=seed)
        B, _ = _rand(64, n, 600, seed=seed + 100)
        assert torch.allclose(dglsp.spspmm(A, B).to_dense(), A.to_dense() @ B.to_dense(), atol=1e-5)


@pytest.mark.parametrize("op", ["add", "sub"])
def test_sparse_add_sub_native_merge(op):
    A, va = _rand(33, 27, 150, torch.float32, requires_grad=True, seed=8)
    B, vb = _rand(33, 27, 170, torch.float32, requires_grad=True, seed=9)
    C = getattr(dglsp, op)(A, B)
    ref = A.to_dense() + B.to_dense() if op == "add" else A.to_dense() - B.to_dense()
    assert torch.allclose(C.to_dense(), ref)
    assert not C.has_duplicate()
    w = torch.randn(33, 27)
    (C.to_dense() * w).sum().backward()
    r, c = A.coo()
    assert torch.allclose(va.grad, w[r.long(), c.long()])
    r, c = B.coo()
    assert torch.allclose(vb.grad, w[r.long(), c.long()] * (1 if op == "add" else -1))


def test_rowdot_gallop_path_matches_merge():
    # one very long row against many short ones exercises the binary-search branch
    n = 4000
    long_cols = torch.arange(0, n, 2)
    xp = torch.tensor([0, long_cols.numel()])
    xv = torch.randn(long_cols.numel())
    g = torch.Generator().manual_seed(0)
    rows_y = 50
    ycols = [torch.sort(torch.randperm(n, generator=g)[:5]).values for _ in range(rows_y)]
    yp = torch.tensor([0] + [5] * rows_y).cumsum(0)
    yi = torch.cat(ycols)
    yv = torch.randn(yi.numel())
    rows = torch.zeros(rows_y, dtype=torch.int64)
    cols = torch.arange(rows_y)
    out = cpu_struct.csr_rowdot(rows, cols, xp, long_cols.contiguous(), xv, yp, yi, yv)
    X = torch.zeros(n)
    X[long_cols] = xv
    Y = torch.zeros(rows_y, n)
    Y[torch.arange(rows_y).repeat_interleave(5), yi] = yv
    assert torch.allclose(out, Y @ X, atol=1e-5)


def test_adj_product_and_sum_graph_use_native_path():
    g = dgl.rand_graph(60, 300)
    g = dgl.to_simple(g)
    g.edata["w"] = torch.rand(g.num_edges(), requires_grad=True)
    p = dgl.adj_product_graph(g, g, "w")
    u, v = g.edges()
    dense = torch.zeros(60, 60).index_put((u.long(), v.long()), g.edata["w"].detach())
    pu, pv = p.edges()
    assert torch.allclose(p.edata["w"], (dense @ dense)[pu.long(), pv.long()], atol=1e-5)
    assert p.num_edges() == int(((dense != 0).float() @ (dense != 0).float() != 0).sum())
    p.edata["w"].sum().backward()
    assert g.edata["w"].grad is not None and torch.isfinite(g.edata["w"].grad).all()
    s = dgl.adj_sum_graph([g, g], "w")
    su, sv = s.edges()
    assert torch.allclose(s.edata["w"], 2 * dense[su.long(), sv.long()], atol=1e-6)


def test_multilevel_partitioner_quality_and_balance():
    """The multilevel k-way partitioner: recovers planted communities, beats the streaming partitioner on a mesh by a wide
    margin, respects the balance constraint (also with edge balancing) and is deterministic for a fixed seed."""
    import dgl_b200 as dgl
    from dgl_b200 import random as dgl_random
    from dgl_b200.partition import metis_partition_assignment

    def cut(g, part):
        u, v = g.edges()
        return (part[u.long()] != part[v.long()]).float().mean().item()

    torch.manual_seed(0)
    n, C = 8000, 8
    comm = torch.arange(n) % C
    src = torch.randint(0, n, (100000,))
    same = torch.rand(100000) < 0.9
    dst = torch.where(same, torch.randint(0, n // C, (100000,)) * C + comm[src], torch.randint(0, n, (100000,)))
    g = dgl.graph((torch.cat([src, dst]), torch.cat([dst, src])), num_nodes=n)
    p = metis_partition_assignment(g, C)
    assert cut(g, p) < 0.1                                            # planted noise floor: 0.1 * 7 / 8
    sizes = torch.bincount(p, minlength=C).float()
    assert sizes.max() / sizes.mean() <= 1.04
    idx = torch.arange(10000).view(100, 100)
    e = torch.cat([torch.stack([idx[:, :-1].reshape(-1), idx[:, 1:].reshape(-1)]),
                   torch.stack([idx[:-1].reshape(-1), idx[1:].reshape(-1)])], 1)
    grid = dgl.graph((torch.cat([e[0], e[1]]), torch.cat([e[1], e[0]])), num_nodes=10000)
    ml = metis_partition_assignment(grid, 4)
    ldg = metis_partition_assignment(grid, 4, algorithm="ldg")
    assert cut(grid, ml) < 0.5 * cut(grid, ldg) and cut(grid, ml) < 0.04
    eb = metis_partition_assignment(g, 4, balance_edges=True)
    deg = g.in_degrees().float()
    load = torch.zeros(4).index_add_(0, eb, deg)
    assert load.max() / load.mean() <= 1.06
    dgl_random.seed(5)
    a = metis_partition_assignment(grid, 4)
    dgl_random.seed(5)
    assert torch.equal(a, metis_partition_assignment(grid, 4))
    assert metis_partition_assignment(grid, 1).unique().tolist() == [0]
    with pytest.raises(dgl.DGLError):
        metis_partition_assignment(grid, 2, algorithm="nope")
