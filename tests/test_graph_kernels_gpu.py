"""Native graph-algorithm kernels (csrc/cuda/graph_kernels.cu) against torch / CPU oracles and structural invariants."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200 import _graph_kernels as gk

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _edge_set(g):
    u, v = g.edges()
    return set(zip(u.tolist(), v.tolist()))


def test_random_walk_uniform_restart_weighted(idtype):
    g = dgl.rand_graph(500, 4000, idtype=idtype).to(DEV)
    es = _edge_set(g)
    seeds = torch.arange(0, 500, device=DEV, dtype=idtype)
    tr, eids, types = dgl.sampling.random_walk(g, seeds, length=6, return_eids=True)
    assert tr.shape == (500, 7) and eids.shape == (500, 6) and types.shape[0] == 7
    t = tr.tolist()
    u, v = g.edges()
    for w in range(500):
        for s in range(6):
            a, b = t[w][s], t[w][s + 1]
            if b < 0:
                assert all(x < 0 for x in t[w][s + 1:])
                break
            assert (a, b) in es
            e = int(eids[w, s])
            assert int(u[e]) == a and int(v[e]) == b
    tr2, _ = dgl.sampling.random_walk(g, seeds, length=4, restart_prob=1.0)
    assert bool((tr2[:, 1:] == -1).all())
    # zero-probability edges are never walked
    p = torch.ones(g.num_edges(), device=DEV)
    p[::2] = 0
    g.edata["p"] = p
    tr3, e3, _ = dgl.sampling.random_walk(g, seeds, length=5, prob="p", return_eids=True)
    valid = e3[e3 >= 0].long()
    assert bool((p[valid] > 0).all())


def test_metapath_walk_hetero():
    g = dgl.heterograph({("user", "buys", "item"): (torch.randint(0, 50, (400,)), torch.randint(0, 30, (400,))),
                         ("item", "bought_by", "user"): (torch.randint(0, 30, (400,)), torch.randint(0, 50, (400,)))}).to(DEV)
    tr, types = dgl.sampling.random_walk(g, torch.arange(50, device=DEV), metapath=["buys", "bought_by"] * 3)
    assert tr.shape == (50, 7) and types.tolist() == [g.get_ntype_id("user"), g.get_ntype_id("item")] * 3 + [g.get_ntype_id("user")]
    ok = tr >= 0
    assert int(tr[:, 1::2][ok[:, 1::2]].max()) < 30 and int(tr[:, 0::2][ok[:, 0::2]].max()) < 50


def test_node2vec_walk_valid_and_biased():
    g = dgl.rand_graph(300, 3000).to(DEV)
    g = dgl.to_bidirected(g.cpu()).to(DEV)
    es = _edge_set(g)
    seeds = torch.arange(300, device=DEV)
    tr = dgl.sampling.node2vec_random_walk(g, seeds, p=1.0, q=1.0, walk_length=8)
    assert tr.shape == (300, 9)
    t = tr.tolist()
    for w in range(300):
        for s in range(8):
            if t[w][s + 1] < 0:
                break
            assert (t[w][s], t[w][s + 1]) in es
    # huge p: returning to the previous vertex is (almost) never accepted when there is an alternative
    trp = dgl.sampling.node2vec_random_walk(g, seeds, p=1e6, q=1.0, walk_length=8)
    back = (trp[:, 2:] == trp[:, :-2]) & (trp[:, 2:] >= 0)
    deg = g.out_degrees()
    multi = deg[trp[:, 1:-1].clamp(min=0).long()] > 1
    assert int((back & multi).sum()) == 0


def test_labor0_native(idtype):
    g = dgl.rand_graph(2000, 60000, idtype=idtype).to(DEV)
    seeds = torch.arange(0, 1000, device=DEV, dtype=idtype)
    sg, w = dgl.sampling.sample_labors(g, seeds, 5, random_seed=7)
    sg2, _ = dgl.sampling.sample_labors(g, seeds, 5, random_seed=7)
    u, v = sg.edges()
    u2, v2 = sg2.edges()
    assert torch.equal(u, u2) and torch.equal(v, v2)                       # counter-based: same seed, same sample
    es = _edge_set(g)
    assert all((a, b) in es for a, b in zip(u.tolist()[:2000], v.tolist()[:2000]))
    assert bool((v.long() < 1000).all())
    indeg = g.in_degrees(seeds).float()
    expected = torch.minimum(indeg, torch.full_like(indeg, 5.0)).sum()
    assert abs(u.shape[0] - float(expected)) < 0.1 * float(expected)
    # a source vertex is either kept or dropped consistently for seeds of equal degree: same variate r_t
    sums = torch.zeros(1000, device=DEV).index_add_(0, v.long(), w[0])
    assert torch.allclose(sums[sums > 0], torch.ones_like(sums[sums > 0]), atol=1e-5)
    full, _ = dgl.sampling.sample_labors(g, seeds, -1)
    assert full.num_edges() == int(indeg.sum())


def test_negative_sampling_native():
    g = dgl.rand_graph(200, 3000).to(DEV)
    u, v = dgl.sampling.global_uniform_negative_sampling(g, 1500)
    es = _edge_set(g)
    pairs = list(zip(u.tolist(), v.tolist()))
    assert len(pairs) > 1000 and all(p not in es for p in pairs) and all(a != b for a, b in pairs)
    assert len(set(pairs)) == len(pairs)


def test_fps_matches_reference_loop():
    pos = torch.randn(4, 3000, 3, device=DEV)
    got = dgl.geometry.farthest_point_sampler(pos, 64, start_idx=5)
    ref = dgl.geometry.farthest_point_sampler(pos.cpu(), 64, start_idx=5)
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("k", [1, 8, 33])
def test_knn_brute_matches_cdist(k):
    x = torch.randn(3000, 3, device=DEV)
    y = torch.randn(700, 3, device=DEV)
    off_d = torch.tensor([0, 1200, 3000], device=DEV)
    off_q = torch.tensor([0, 300, 700], device=DEV)
    idx, dist = gk.knn_brute(x, off_d, y, off_q, k, return_dist=True)
    for s in range(2):
        xd = x[int(off_d[s]):int(off_d[s + 1])]
        yq = y[int(off_q[s]):int(off_q[s + 1])]
        d = torch.cdist(yq, xd) ** 2
        rd, ri = torch.topk(d, k, dim=1, largest=False)
        assert torch.allclose(dist[int(off_q[s]):int(off_q[s + 1])], rd, atol=1e-4)
        same = (idx[int(off_q[s]):int(off_q[s + 1])] - int(off_d[s])) == ri
        assert float(same.float().mean()) > 0.999
    g = dgl.knn_graph(x[:500], 6)
    assert g.num_edges() == 3000 and bool((g.in_degrees() == 6).all())


def test_neighbor_matching_native():
    g = dgl.to_bidirected(dgl.rand_graph(1000, 6000)).to(DEV)
    c = dgl.geometry.neighbor_matching(g, relabel_idx=False)
    es = _edge_set(g)
    cl = c.tolist()
    groups = {}
    for i, r in enumerate(cl):
        groups.setdefault(r, []).append(i)
    assert all(len(m) <= 2 for m in groups.values())
    assert all(len(m) == 1 or (m[0], m[1]) in es for m in groups.values())
    assert sum(len(m) == 2 for m in groups.values()) > 300           # a maximal matching pairs most vertices
    c2 = dgl.geometry.neighbor_matching(g, e_weights=torch.rand(g.num_edges(), device=DEV))
    assert int(c2.max()) + 1 == len(set(c2.tolist()))


@pytest.mark.parametrize("mode", ["remainder", "range"])
def test_partition_ops_native(mode, idtype):
    part = dgl.partition.NDArrayPartition(10000, 7, mode=mode)
    idx = torch.randint(0, 10000, (5000,), dtype=idtype)
    perm_ref, cnt_ref = part.generate_permutation(idx)
    perm, cnt = part.generate_permutation(idx.to(DEV))
    assert torch.equal(perm.cpu(), perm_ref) and torch.equal(cnt.cpu(), cnt_ref)
    own, _, loc = gk.partition_owner(idx.to(DEV), 7, part._ranges if mode == "range" else None, want_local=True)
    assert torch.equal(own.cpu().long(), part.owner(idx).long()) and torch.equal(loc.cpu(), part.map_to_local(idx))


def test_structure_helpers_native(idtype):
    from dgl_b200.graph_index import coo_to_csr
    row = torch.randint(0, 300, (5000,), dtype=idtype)
    col = torch.randint(0, 300, (5000,), dtype=idtype)
    ip_c, ix_c, e_c = coo_to_csr(row, col, 300, idtype)
    ip_g, ix_g, e_g = coo_to_csr(row.to(DEV), col.to(DEV), 300, idtype)
    assert torch.equal(ip_g.cpu(), ip_c) and torch.equal(ix_g.cpu(), ix_c) and torch.equal(e_g.cpu(), e_c)
    rows = gk.expand_indptr(ip_g)
    assert torch.equal(rows.cpu(), torch.sort(row)[0])
    ids = torch.arange(300, device=DEV, dtype=idtype) * 3
    assert torch.equal(gk.expand_indptr(ip_g, node_ids=ids).cpu(), torch.sort(row)[0] * 3)


def test_cooperative_rank_sort_native():
    from dgl_b200.graphbolt.impl.cooperative import rank_sort
    ids = torch.randint(0, 100000, (5000,))
    s_c, p_c, o_c = rank_sort(ids, 0, 8, seed=3)
    s_g, p_g, o_g = rank_sort(ids.to(DEV), 0, 8, seed=3)
    assert torch.equal(s_g.cpu(), s_c) and torch.equal(p_g.cpu(), p_c) and torch.equal(o_g.cpu(), o_c)


# ---------------------------------------------------------------------------------------------------------------------
# structure construction kernels (csrc/cuda/structure_build.cu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("idtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("n,m", [(1, 5), (1000, 50000), (70000, 2000000)])
def test_native_coo_to_csr_and_transpose(idtype, n, m):
    from dgl_b200.ops.cuda_struct import cuda_coo_to_csr, cuda_csr_transpose
    g = torch.Generator().manual_seed(n + m)
    row = torch.randint(0, n, (m,), generator=g).to(idtype).cuda()
    col = torch.randint(0, n + 3, (m,), generator=g).to(idtype).cuda()
    indptr, indices, eids = cuda_coo_to_csr(row, col, n)
    torch.cuda.synchronize()
    # reference: stable sort by row
    perm = torch.sort(row.long(), stable=True)[1]
    assert torch.equal(eids.long(), perm)
    assert torch.equal(indices, col[perm])
    counts = torch.bincount(row.long(), minlength=n)
    assert torch.equal(indptr.long(), torch.cat([counts.new_zeros(1), counts.cumsum(0)]))
    tp, ti, te = cuda_csr_transpose(indptr, indices, n, n + 3)
    torch.cuda.synchronize()
    srow = row[perm]
    perm2 = torch.sort(indices.long(), stable=True)[1]
    assert torch.equal(te.long(), perm2) and torch.equal(ti, srow[perm2])
    c2 = torch.bincount(indices.long(), minlength=n + 3)
    assert torch.equal(tp.long(), torch.cat([c2.new_zeros(1), c2.cumsum(0)]))


def test_native_coalesce_and_graph_formats():
    from dgl_b200.ops.cuda_struct import cuda_coalesce
    g = torch.Generator().manual_seed(0)
    n, m = 500, 40000
    row = torch.randint(0, n, (m,), generator=g).int().cuda()
    col = torch.randint(0, 60, (m,), generator=g).int().cuda()
    val = torch.randn(m, 3, generator=g).cuda()
    r, c, v, first = cuda_coalesce(row, col, val, n, 60, want_first=True)
    dense = torch.zeros(n, 60, 3, device="cuda").index_put_((row.long(), col.long()), val, accumulate=True)
    key = r.long() * 60 + c.long()
    assert bool((key[1:] > key[:-1]).all())                                  # sorted, distinct
    assert torch.allclose(dense[r.long(), c.long()], v, atol=1e-4)
    assert int((dense.abs().sum(-1) > 0).sum()) <= r.shape[0]
    assert torch.equal(row[first], r) and torch.equal(col[first], c)
    # the graph index builds its formats with the native kernels on CUDA
    import dgl_b200 as dgl
    u = torch.randint(0, 3000, (100000,), generator=g)
    w = torch.randint(0, 3000, (100000,), generator=g)
    gg = dgl.graph((u, w), num_nodes=3000).to("cuda").formats(["coo", "csr", "csc"])
    gg.create_formats_()
    gc = dgl.graph((u, w), num_nodes=3000)
    assert torch.equal(gg.in_degrees().cpu(), gc.in_degrees()) and torch.equal(gg.out_degrees().cpu(), gc.out_degrees())
    x = torch.randn(3000, 8, generator=g)
    assert torch.allclose(dgl.ops.copy_u_sum(gg, x.cuda()).cpu(), dgl.ops.copy_u_sum(gc, x), atol=1e-4)


@pytest.mark.parametrize("shape,nnz", [((300, 200, 250), 3000), ((5000, 4000, 3000), 60000), ((50, 50, 50), 2000)])
def test_native_spgemm_matches_cpu_gustavson(shape, nnz):
    import dgl_b200.sparse as dglsp
    g = torch.Generator().manual_seed(nnz)
    M, K, N = shape
    ai = torch.stack([torch.randint(0, M, (nnz,), generator=g), torch.randint(0, K, (nnz,), generator=g)])
    bi = torch.stack([torch.randint(0, K, (nnz,), generator=g), torch.randint(0, N, (nnz,), generator=g)])
    av, bv = torch.randn(nnz, generator=g), torch.randn(nnz, generator=g)
    A, B = dglsp.spmatrix(ai, av, (M, K)).coalesce(), dglsp.spmatrix(bi, bv, (K, N)).coalesce()
    C_cpu = dglsp.spspmm(A, B)
    Ag = dglsp.spmatrix(A.indices().cuda(), A.val.cuda(), (M, K))
    Bg = dglsp.spmatrix(B.indices().cuda(), B.val.cuda(), (K, N))
    C_gpu = dglsp.spspmm(Ag, Bg)
    assert C_gpu.nnz == C_cpu.nnz
    assert torch.allclose(C_gpu.to_dense().cpu(), C_cpu.to_dense(), atol=1e-3)
