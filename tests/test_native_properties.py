"""Property-based tests (hypothesis) of the host native kernels: random shapes, densities, id widths and degenerate
inputs against dense / brute-force oracles."""
import numpy as np
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import dgl_b200 as dgl
import dgl_b200.sparse as dglsp
from dgl_b200.ops import cpu_struct
from dgl_b200.transforms.functional import kdtree_knn, rcmk_perm

SET = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])


def _mat(seed, m, n, density, idtype):
    g = torch.Generator().manual_seed(seed)
    mask = torch.rand(m, n, generator=g) < density
    r, c = mask.nonzero(as_tuple=True)
    p = torch.randperm(r.numel(), generator=g)
    vals = torch.randn(r.numel(), generator=g, dtype=torch.float64)
    return dglsp.from_coo(r[p].to(idtype), c[p].to(idtype), vals[p], (m, n))


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), m=st.integers(1, 24), k=st.integers(1, 24), n=st.integers(1, 24),
       da=st.floats(0.0, 0.6), db=st.floats(0.0, 0.6), wide=st.booleans())
def test_spgemm_and_sum_match_dense(seed, m, k, n, da, db, wide):
    idt = torch.int64 if wide else torch.int32
    A, B = _mat(seed, m, k, da, idt), _mat(seed + 1, k, n, db, idt)
    C = dglsp.spspmm(A, B)
    assert torch.allclose(C.to_dense(), A.to_dense() @ B.to_dense(), atol=1e-10)
    ip, idx, _ = C.csr()
    for r in range(m):
        row = idx[ip[r]:ip[r + 1]]
        assert bool((row[1:] > row[:-1]).all())
    A2 = _mat(seed + 2, m, k, db, idt)
    assert torch.allclose((A + A2).to_dense(), A.to_dense() + A2.to_dense(), atol=1e-12)
    assert torch.allclose((A - A2).to_dense(), A.to_dense() - A2.to_dense(), atol=1e-12)
    assert torch.allclose((A * A2).to_dense(), A.to_dense() * A2.to_dense(), atol=1e-12)


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 20), d=st.floats(0.05, 0.7))
def test_rowdot_equals_dense_inner_products(seed, n, d):
    X, Y = _mat(seed, n, n + 3, d, torch.int64), _mat(seed + 7, n, n + 3, d, torch.int64)
    xp, xi, xe = X.csr()
    yp, yi, ye = Y.csr()
    xi, xperm = cpu_struct.sort_rows(xp, xi, xe)
    yi, yperm = cpu_struct.sort_rows(yp, yi, ye)
    g = torch.Generator().manual_seed(seed)
    rows = torch.randint(0, n, (30,), generator=g)
    cols = torch.randint(0, n, (30,), generator=g)
    out = cpu_struct.csr_rowdot(rows, cols, xp, xi, X.val[xperm].contiguous(), yp, yi, Y.val[yperm].contiguous())
    ref = (X.to_dense()[rows] * Y.to_dense()[cols]).sum(1)
    assert torch.allclose(out, ref, atol=1e-10)


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 400), q=st.integers(1, 50), dim=st.integers(1, 6), k=st.integers(1, 12),
       dup=st.booleans())
def test_kdtree_distances_match_bruteforce(seed, n, q, dim, k, dup):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, dim, generator=g, dtype=torch.float64)
    if dup and n > 4:
        x[n // 2:] = x[: n - n // 2].clone()              # many exact duplicates
    y = torch.randn(q, dim, generator=g, dtype=torch.float64)
    idx, d = kdtree_knn(x, y, k, return_distance=True)
    kk = min(k, n)
    ref = torch.topk(torch.cdist(y, x, compute_mode="donot_use_mm_for_euclid_dist"), kk, dim=1, largest=False)[0]
    assert torch.allclose(d[:, :kk], ref, atol=1e-9)
    assert bool((idx[:, kk:] == -1).all()) and bool((idx[:, :kk] >= 0).all())
    # no point is returned twice for one query
    for row in idx[:, :kk].tolist():
        assert len(set(row)) == kk


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), pop=st.integers(1, 3000), frac=st.floats(0.0, 1.0), weighted=st.booleans())
def test_choice_without_replacement_is_a_subset(seed, pop, frac, weighted):
    dgl.seed(seed)
    num = int(pop * frac)
    prob = None
    if weighted:
        prob = torch.rand(pop, generator=torch.Generator().manual_seed(seed)) + 0.01
    out = dgl.random.choice(pop, num, replace=False, prob=prob)
    assert out.numel() == num and out.unique().numel() == num
    assert num == 0 or (int(out.min()) >= 0 and int(out.max()) < pop)


@settings(**SET)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 120), e=st.integers(0, 400), wide=st.booleans())
def test_traversal_and_rcm_invariants(seed, n, e, wide):
    g0 = torch.Generator().manual_seed(seed)
    u, v = torch.randint(0, n, (e,), generator=g0), torch.randint(0, n, (e,), generator=g0)
    g = dgl.graph((u, v), num_nodes=n, idtype=torch.int64 if wide else torch.int32)
    # BFS levels: every node appears once, and each non-source node has a predecessor in the previous level
    levels = dgl.bfs_nodes_generator(g, 0)
    flat = torch.cat(levels).long()
    assert flat.unique().numel() == flat.numel()
    depth = torch.full((n,), -1)
    for i, l in enumerate(levels):
        depth[l.long()] = i
    reached = depth >= 0
    if e:
        uu, vv = u[reached[u]], v[reached[u]]
        assert bool(reached[vv].all()) and bool((depth[vv] <= depth[uu] + 1).all())
    perm = rcmk_perm(g)
    assert sorted(perm.tolist()) == list(range(n))
    # DFS tree edges discover each reachable node exactly once
    if g.out_degrees(0) > 0:
        tree = torch.cat(dgl.dfs_edges_generator(g, 0)).long()
        heads = g.edges()[1][tree].long()
        assert heads.unique().numel() == heads.numel() and int(reached.sum()) == heads.numel() + 1 - int((heads == 0).any())
