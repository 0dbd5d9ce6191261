"""Host geometry natives (csrc/cpu/spatial.cc): kd-tree kNN, farthest point sampling, neighbour matching.
Oracles: dense torch.cdist top-k, the textbook FPS loop, matching validity (reference
tests/python/common/transforms/test_transform.py::test_knn_cpu, tests/python/pytorch/geometry/test_geometry.py)."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200.geometry import farthest_point_sampler, neighbor_matching
from dgl_b200.transforms.functional import kdtree_knn


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dim", [1, 3, 8])
def test_kdtree_matches_bruteforce(dtype, dim):
    g = torch.Generator().manual_seed(dim)
    x = torch.randn(3000, dim, generator=g, dtype=dtype)
    y = torch.randn(257, dim, generator=g, dtype=dtype)
    idx, d = kdtree_knn(x, y, 7, return_distance=True)
    rd, _ = torch.topk(torch.cdist(y, x, compute_mode="donot_use_mm_for_euclid_dist"), 7, dim=1, largest=False)
    assert torch.allclose(d, rd, atol=1e-4 if dtype == torch.float32 else 1e-10)
    # returned indices really are at the returned distances
    assert torch.allclose((x[idx] - y[:, None]).norm(dim=-1), d, atol=1e-4)


def test_kdtree_degenerate_clouds():
    x = torch.zeros(600, 3)
    x[:300, 1] = 2.0
    idx, d = kdtree_knn(x, x[[0, 599]], 5, return_distance=True)
    assert torch.all(d == 0) and torch.all(idx[0] < 300) and torch.all(idx[1] >= 300)
    assert len(set(idx[0].tolist())) == 5
    # fewer points than k: padded with -1 / inf
    idx, d = kdtree_knn(torch.randn(3, 2), torch.randn(2, 2), 5, return_distance=True)
    assert torch.all(idx[:, 3:] == -1) and torch.isinf(d[:, 3:]).all() and torch.all(idx[:, :3] >= 0)


def test_segmented_kdtree_stays_inside_segments():
    x = torch.randn(1000, 3)
    offs = [0, 400, 450, 1000]
    idx, _ = kdtree_knn(x, x, 6, offs, offs)
    for s in range(3):
        seg = idx[offs[s]:offs[s + 1]]
        assert torch.all((seg >= offs[s]) & (seg < offs[s + 1]))
    assert torch.all(idx[:, 0] == torch.arange(1000))  # nearest neighbour of a point is itself


@pytest.mark.parametrize("exclude_self", [False, True])
@pytest.mark.parametrize("dist", ["euclidean", "cosine"])
def test_knn_graph_cpu_uses_tree_and_agrees_with_dense(exclude_self, dist):
    x = torch.randn(500, 3)
    g = dgl.knn_graph(x, 4, dist=dist, exclude_self=exclude_self)
    assert g.num_edges() == 500 * 4
    u, v = g.edges()
    assert torch.all(g.in_degrees() == 4)
    pts = torch.nn.functional.normalize(x, dim=-1) if dist == "cosine" else x
    D = torch.cdist(pts, pts, compute_mode="donot_use_mm_for_euclid_dist")
    if exclude_self:
        D.fill_diagonal_(float("inf"))
        assert not torch.any(u == v)
    kth = torch.topk(D, 4, dim=1, largest=False)[0][:, -1]
    assert torch.all(D[v, u] <= kth[v] + 1e-5)
    sg = dgl.segmented_knn_graph(x, 4, [200, 300], dist=dist, exclude_self=exclude_self)
    su, sv = sg.edges()
    assert sg.num_edges() == 2000 and torch.all((su < 200) == (sv < 200))


def _fps_ref(pos, n, start):
    B, N, _ = pos.shape
    dist = torch.full((B, N), float("inf"), dtype=pos.dtype)
    cur = torch.full((B,), start)
    out = torch.empty(B, n, dtype=torch.long)
    ar = torch.arange(B)
    for i in range(n):
        out[:, i] = cur
        dist = torch.minimum(dist, ((pos - pos[ar, cur].unsqueeze(1)) ** 2).sum(-1))
        cur = dist.argmax(1)
    return out


@pytest.mark.parametrize("shape", [(5, 300, 3), (1, 40000, 3), (64, 50, 2)])
def test_fps_native_equals_reference_loop(shape):
    pos = torch.randn(*shape)
    out = farthest_point_sampler(pos, 20, start_idx=1)
    assert torch.equal(out, _fps_ref(pos, 20, 1))
    rnd = farthest_point_sampler(pos.double(), 10)
    assert rnd.shape == (shape[0], 10) and all(len(set(r.tolist())) == 10 for r in rnd)


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("idtype", [torch.int32, torch.int64])
def test_neighbor_matching_is_a_maximal_matching(weighted, idtype):
    g = dgl.rand_graph(400, 1600).astype(idtype)
    w = torch.rand(1600) if weighted else None
    torch.manual_seed(7)
    m = neighbor_matching(g, w, relabel_idx=False)
    assert m.dtype == idtype
    u, v = [t.long() for t in g.edges()]
    adj = torch.zeros(400, 400, dtype=torch.bool)
    adj[u, v] = adj[v, u] = True
    sizes = torch.bincount(m.long(), minlength=400)
    assert sizes.max() <= 2
    for c in torch.nonzero(sizes == 2).view(-1).tolist():
        a, b = torch.nonzero(m == c).view(-1).tolist()
        assert adj[a, b] and c == min(a, b)
    # maximal: no edge joins two singletons
    single = sizes[m.long()] == 1
    nz = u != v
    assert not torch.any(single[u[nz]] & single[v[nz]])
    torch.manual_seed(7)  # same visiting order -> same matching, now relabelled to 0..C-1
    r = neighbor_matching(g, w)
    assert int(r.max()) + 1 == int((sizes > 0).sum())


def test_nn_descent_recall_segments_and_knn_entry():
    """Approximate kNN: neighbours stay inside their segment, no duplicates, self first, high recall on clustered data;
    ``knn(..., algorithm='nn-descent')`` switches to it for large sets."""
    import dgl_b200 as dgl
    from dgl_b200.transforms import functional as TF
    from dgl_b200.transforms.nn_descent import nn_descent
    torch.manual_seed(0)
    centres = torch.randn(40, 24) * 4
    x = centres[torch.randint(0, 40, (5000,))] + torch.randn(5000, 24)
    offs = [0, 1800, 5000]
    k = 8
    nb = nn_descent(x, offs, k)
    hits = 0
    for lo, hi in zip(offs[:-1], offs[1:]):
        got = nb[lo:hi]
        assert (got >= lo).all() and (got < hi).all()
        exact = torch.cdist(x[lo:hi], x[lo:hi]).topk(k, largest=False).indices + lo
        hits += sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(exact, got))
    assert hits / (5000 * k) > 0.93
    assert (nb[:, 0] == torch.arange(5000)).all() and all(len(set(r)) == k for r in nb.tolist())
    # tiny segments (fewer points than k + 1) are still served
    small = nn_descent(torch.randn(7, 3), [0, 3, 7], 3)
    assert small.shape == (7, 3) and (small[:3] < 3).all() and (small[3:] >= 3).all()
    old = TF.NN_DESCENT_MIN_POINTS
    TF.NN_DESCENT_MIN_POINTS = 1000
    try:
        pairs = dgl.knn(k, x, [1800, 3200], algorithm="nn-descent")
    finally:
        TF.NN_DESCENT_MIN_POINTS = old
    assert pairs.shape == (2, 5000 * k) and pairs[0].tolist() == torch.arange(5000).repeat_interleave(k).tolist()
    g = dgl.knn_graph(x, k, algorithm="nn-descent")
    assert g.num_edges() == 5000 * k
