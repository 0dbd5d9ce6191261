"""sm_100a kernels vs plain PyTorch fp32/fp64 references (run on the B200 box: pytest -m gpu)."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200 import ops
from graph_cases import get_cases
from test_ops_cpu import SHAPES, oracle_sddmm, oracle_spmm

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tol(dtype):
    return {torch.float32: (1e-4, 1e-4), torch.float64: (1e-9, 1e-9), torch.float16: (2e-2, 2e-2),
            torch.bfloat16: (6e-2, 6e-2)}[dtype]


@pytest.mark.parametrize("case", get_cases(["homo", "bipartite"]), ids=lambda f: f.__name__)
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "copy_lhs", "copy_rhs"])
@pytest.mark.parametrize("red", ["sum", "max", "min", "mean"])
@pytest.mark.parametrize("shp", SHAPES[:4], ids=str)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gspmm_cuda(case, op, red, shp, dtype, idtype):
    g = case(idtype)
    gg = g.to(DEV)
    torch.manual_seed(0)
    x = (torch.rand((g.num_src_nodes(),) + shp[0], dtype=dtype) + 0.5)
    w = (torch.rand((g.num_edges(),) + shp[1], dtype=dtype) + 0.5)
    xc, wc = x.clone().requires_grad_(), w.clone().requires_grad_()
    xg, wg = x.to(DEV).requires_grad_(), w.to(DEV).requires_grad_()
    out = ops.gspmm(gg, op, red, xg, wg)
    ref = oracle_spmm(g, op, red, xc, wc)
    rt, at = _tol(dtype)
    assert torch.allclose(out.cpu(), ref, rtol=rt, atol=at)
    go = torch.rand_like(ref)
    need_c = [t for t, used in ((xc, op != "copy_rhs"), (wc, op != "copy_lhs")) if used]
    need_g = [t for t, used in ((xg, op != "copy_rhs"), (wg, op != "copy_lhs")) if used]
    g1 = torch.autograd.grad(out, need_g, go.to(DEV))
    g2 = torch.autograd.grad(ref, need_c, go)
    for a, b in zip(g1, g2):
        assert torch.allclose(a.cpu(), b, rtol=rt * 10, atol=at * 10)


@pytest.mark.parametrize("F", [4, 32, 100, 128, 256, 602])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("weighted", [False, True])
def test_spmm_vec_path(F, dtype, weighted, idtype):
    torch.manual_seed(1)
    n, E = 3000, 40000
    u = torch.randint(0, n, (E,))
    v = torch.randint(0, n, (E,))
    v[:5000] = 7  # one heavy row
    g = dgl.graph((u, v), num_nodes=n, idtype=idtype).to(DEV)
    x = torch.randn(n, F, device=DEV).to(dtype)
    w = torch.rand(E, 1, device=DEV).to(dtype) if weighted else None
    out = ops.gspmm(g, "mul" if weighted else "copy_lhs", "sum", x, w)
    m = x.float()[u.to(DEV)] * (w.float() if weighted else 1.0)
    ref = torch.zeros(n, F, device=DEV).index_add_(0, v.to(DEV), m)
    rt, at = _tol(dtype)
    scale = ref.abs().max().item()
    assert (out.float() - ref).abs().max().item() <= at * max(scale, 1.0) + 1e-3


@pytest.mark.parametrize("case", get_cases(["homo", "bipartite"]), ids=lambda f: f.__name__)
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "dot", "copy_lhs", "copy_rhs"])
@pytest.mark.parametrize("lt,rt", [("u", "v"), ("u", "e"), ("e", "v"), ("v", "u")])
@pytest.mark.parametrize("shp", [((5,), (5,)), ((4, 1), (4, 3)), ((2, 3), (2, 3))], ids=str)
def test_gsddmm_cuda(case, op, lt, rt, shp, idtype):
    g = case(idtype)
    gg = g.to(DEV)
    torch.manual_seed(1)
    n = {"u": g.num_src_nodes(), "v": g.num_dst_nodes(), "e": g.num_edges()}
    s0, s1 = shp
    if op == "dot":
        s0, s1 = s0[:-1] + (40,), s1[:-1] + (40,)
    x = torch.rand((n[lt],) + s0) + 0.5
    y = torch.rand((n[rt],) + s1) + 0.5
    xc, yc = x.clone().requires_grad_(), y.clone().requires_grad_()
    xg, yg = x.to(DEV).requires_grad_(), y.to(DEV).requires_grad_()
    out = ops.gsddmm(gg, op, xg, yg, lt, rt)
    ref = oracle_sddmm(g, op, xc, yc, lt, rt)
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-4)
    go = torch.rand_like(ref)
    need_c = [t for t, used in ((xc, op != "copy_rhs"), (yc, op != "copy_lhs")) if used]
    need_g = [t for t, used in ((xg, op != "copy_rhs"), (yg, op != "copy_lhs")) if used]
    g1 = torch.autograd.grad(out, need_g, go.to(DEV))
    g2 = torch.autograd.grad(ref, need_c, go)
    for a, b in zip(g1, g2):
        assert torch.allclose(a.cpu(), b, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("norm_by", ["dst", "src"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_edge_softmax_cuda(norm_by, dtype, idtype):
    g = get_cases(["random"])[0](idtype)
    gg = g.to(DEV)
    s = torch.randn(g.num_edges(), 8, 1)
    sc = s.clone().requires_grad_()
    sg = s.to(DEV).to(dtype).requires_grad_()
    out = ops.edge_softmax(gg, sg, norm_by=norm_by)
    ref = ops.edge_softmax(g, sc, norm_by=norm_by)
    rt, at = _tol(dtype)
    assert torch.allclose(out.float().cpu(), ref, rtol=rt, atol=at)
    go = torch.rand_like(ref)
    (g1,) = torch.autograd.grad(out, sg, go.to(DEV).to(dtype))
    (g2,) = torch.autograd.grad(ref, sc, go)
    assert torch.allclose(g1.float().cpu(), g2, rtol=rt * 5, atol=at * 5)


@pytest.mark.parametrize("permute", [False, True])
@pytest.mark.parametrize("heads", [1, 3, 8, 40])
def test_edge_softmax_hub_rows(heads, permute, monkeypatch):
    """Hub destinations (> 2048 in-edges) go through the block-per-row kernel; ordinary rows stay warp-per-row.
    ``permute`` forces the large-graph path that brings the scores into CSC order first."""
    import dgl_b200._sparse_ops as so
    monkeypatch.setattr(so, "_PERMUTE_MIN_EDGES", 0 if permute else 1 << 40)
    torch.manual_seed(heads)
    n = 400
    src = torch.cat([torch.randint(0, n, (6000,)), torch.randint(0, n, (3000,)), torch.randint(0, n, (4000,))])
    dst = torch.cat([torch.full((6000,), 7), torch.full((3000,), 100), torch.randint(0, n, (4000,))])
    perm = torch.randperm(src.numel())  # shuffled edge ids: CSC carries a real eid indirection
    g = dgl.graph((src[perm], dst[perm]), num_nodes=n)
    gg = g.to(DEV)
    s = torch.randn(g.num_edges(), heads) * 3
    sc = s.clone().requires_grad_()
    sg = s.to(DEV).requires_grad_()
    out = ops.edge_softmax(gg, sg)
    ref = ops.edge_softmax(g, sc)
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=1e-6)
    go = torch.rand_like(ref)
    (g1,) = torch.autograd.grad(out, sg, go.to(DEV))
    (g2,) = torch.autograd.grad(ref, sc, go)
    assert torch.allclose(g1.cpu(), g2, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 32), (4, 16), (1, 128), (2, 24), (1, 5)])
def test_u_dot_v_vectorised(dtype, shape):
    """u_dot_v through the 16-byte-vector kernel (and the scalar fallback for odd widths) vs the CPU kernels."""
    torch.manual_seed(0)
    g = dgl.rand_graph(300, 5000)
    gg = g.to(DEV)
    x = torch.randn(300, *shape)
    y = torch.randn(300, *shape)
    ref = ops.u_dot_v(g, x, y)
    out = ops.u_dot_v(gg, x.to(DEV).to(dtype), y.to(DEV).to(dtype))
    tol = 1e-4 if dtype == torch.float32 else 0.15
    assert torch.allclose(out.float().cpu(), ref, rtol=tol, atol=tol * 4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(1, 128), (1, 602), (1, 100), (1, 2), (1, 1030), (8, 64), (3, 8), (1, 6)])
@pytest.mark.parametrize("csc_first", [False, True])
def test_u_dot_v_staged_kernel(dtype, shape, csc_first, idtype):
    """Staged dot kernel: 16 / 8 / 4-byte vector widths, long rows, multi-head, edge count not a multiple of 32, skewed
    degrees; with the CSC present the kernel walks the edges in destination order and scatters through the edge-id map -
    the result must be in edge-id order either way. Oracle: dense fp32 torch."""
    torch.manual_seed(1)
    n = 257
    src = torch.cat([torch.randint(0, n, (4001,)), torch.randint(0, n, (1500,))])
    dst = torch.cat([torch.randint(0, n, (4001,)), torch.full((1500,), 11)])
    perm = torch.randperm(src.numel())
    src, dst = src[perm], dst[perm]
    gg = dgl.graph((src, dst), num_nodes=n, idtype=idtype, device=DEV)
    if csc_first:
        gg.create_formats_()
    x = torch.randn(n, *shape).to(dtype)
    y = torch.randn(n, *shape).to(dtype)
    want = (x.float()[src] * y.float()[dst]).sum(-1, keepdim=True)
    got = ops.u_dot_v(gg, x.to(DEV), y.to(DEV))
    assert got.shape == want.shape and got.dtype == dtype
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    scale = want.abs().max().item()
    assert (got.float().cpu() - want).abs().max().item() <= tol * max(scale, 1.0)
    # v_dot_u swaps the operand roles but not the edge order
    got2 = ops.v_dot_u(gg, y.to(DEV), x.to(DEV))
    assert (got2.float().cpu() - want).abs().max().item() <= tol * max(scale, 1.0)


def test_u_dot_v_staged_kernel_grad():
    torch.manual_seed(2)
    n = 90
    g = dgl.rand_graph(n, 2000).to(DEV)
    g.create_formats_()
    x = torch.randn(n, 4, 16, device=DEV, requires_grad=True)
    y = torch.randn(n, 4, 16, device=DEV, requires_grad=True)
    out = ops.u_dot_v(g, x, y)
    u, v = g.edges()
    want = (x[u.long()] * y[v.long()]).sum(-1, keepdim=True)
    assert torch.allclose(out, want, atol=1e-4)
    go = torch.randn_like(out)
    gx, gy = torch.autograd.grad(out, (x, y), go)
    wx, wy = torch.autograd.grad(want, (x, y), go)
    assert torch.allclose(gx, wx, atol=1e-3) and torch.allclose(gy, wy, atol=1e-3)


@pytest.mark.parametrize("red", ["sum", "max", "min", "mean"])
def test_segment_reduce_cuda(red):
    seglen = torch.tensor([3, 0, 4, 1, 0, 120])
    x = torch.randn(128, 37)
    a = ops.segment_reduce(seglen, x, red)
    b = ops.segment_reduce(seglen.to(DEV), x.to(DEV), red)
    assert torch.allclose(a, b.cpu(), atol=1e-5)


def test_unique_compact_and_block(idtype):
    torch.manual_seed(3)
    first = torch.randperm(5000)[:300].to(idtype)
    ids = torch.randint(0, 5000, (20000,)).to(idtype)
    from dgl_b200.transforms import unique_and_compact
    u_c, l_c = unique_and_compact(first, ids)
    u_g, l_g = unique_and_compact(first.to(DEV), ids.to(DEV))
    assert torch.equal(u_c, u_g.cpu())
    assert torch.equal(l_c, l_g.cpu())
    assert torch.equal(u_g[l_g.long()], ids.to(DEV))


def test_sampling_cuda_and_uva(idtype):
    g = dgl.rand_graph(5000, 200000, idtype=idtype)
    g.edata["p"] = torch.rand(g.num_edges())
    seeds = torch.arange(0, 2000, dtype=idtype)
    gg = g.to(DEV)
    for kw in ({}, {"replace": True}, {"prob": "p"}, {"prob": "p", "replace": True}):
        sg = dgl.sampling.sample_neighbors(gg, seeds.to(DEV), 7, **kw)
        u, v = sg.edges()
        assert bool(g.has_edges_between(u.cpu(), v.cpu()).all())
        deg = sg.in_degrees()
        assert int(deg.max()) <= 7
        eid = sg.edata[dgl.EID].cpu()
        eu, ev = g.find_edges(eid)
        assert torch.equal(eu, u.cpu()) and torch.equal(ev, v.cpu())
        if not kw.get("replace"):
            exp = torch.minimum(g.in_degrees()[:2000], torch.tensor(7))
            if "prob" not in kw:
                assert torch.equal(deg[:2000].cpu().long(), exp.long())
            # no duplicates
            assert torch.unique(eid).numel() == eid.numel()
    # UVA: graph pinned on host, seeds on GPU
    g.create_formats_()
    g.pin_memory_()
    sg = dgl.sampling.sample_neighbors(g, seeds.to(DEV), 5)
    assert sg.device.type == "cuda"
    u, v = sg.edges()
    assert bool(g.has_edges_between(u.cpu(), v.cpu()).all())
    g.unpin_memory_()


def test_uva_gather():
    from dgl_b200.utils import gather_pinned_tensor_rows, pin_memory_inplace, unpin_memory_inplace
    for F, dt in ((100, torch.float32), (128, torch.bfloat16), (7, torch.float32), (602, torch.float16)):
        x = torch.randn(10000, F).to(dt)
        pin_memory_inplace(x)
        idx = torch.randint(0, 10000, (5000,), device=DEV)
        out = gather_pinned_tensor_rows(x, idx)
        assert torch.equal(out.cpu(), x[idx.cpu()])
        unpin_memory_inplace(x)
    xd = torch.randn(1000, 33, device=DEV)
    from dgl_b200.ops.index import index_select
    idx = torch.randint(0, 1000, (777,), device=DEV, dtype=torch.int32)
    assert torch.equal(index_select(xd, idx), xd[idx.long()])


def test_minibatch_sage_gpu():
    import torch.nn.functional as F
    from dgl_b200.dataloading import DataLoader, NeighborSampler
    torch.manual_seed(0)
    g = dgl.rand_graph(3000, 60000)
    g.ndata["feat"] = torch.randn(3000, 16)
    g.ndata["label"] = (g.ndata["feat"][:, 0] > 0).long()
    g = g.to(DEV)
    sampler = NeighborSampler([5, 5])
    dl = DataLoader(g, torch.arange(2000, device=DEV), sampler, batch_size=256, shuffle=True)
    l1 = dgl.nn.SAGEConv(16, 32, "mean").to(DEV)
    l2 = dgl.nn.SAGEConv(32, 2, "mean").to(DEV)
    opt = torch.optim.Adam(list(l1.parameters()) + list(l2.parameters()), 1e-2)
    losses = []
    for ep in range(4):
        for inp, out, blocks in dl:
            x = g.ndata["feat"][inp.long()]
            y = g.ndata["label"][out.long()]
            h = l2(blocks[1], F.relu(l1(blocks[0], x)))
            loss = F.cross_entropy(h, y)
            opt.zero_grad(); loss.backward(); opt.step()
            losses.append(loss.item())
    assert losses[-1] < losses[0]
