"""Execution plans of layers that differ from the reference's formulation must stay numerically equivalent:
RelGraphConv aggregate-first (basis regulariser) and the destination-sorted edge order of attention layers."""
import pytest
import torch

import dgl_b200 as dgl


def _rgcn_message_plan(conv, g, x, et, norm):
    lin = conv.linear_r
    W = lin.get_weight()
    u, v = g.edges()
    m = torch.bmm(x[u.long()].unsqueeze(1), W[et.long()]).squeeze(1)
    if norm is not None:
        m = m * norm
    h = torch.zeros(g.num_nodes(), W.shape[2], device=x.device, dtype=x.dtype).index_add_(0, v.long(), m)
    return h + conv.h_bias + x @ conv.loop_weight


@pytest.mark.parametrize("with_norm", [False, True])
def test_relgraphconv_aggregate_first_matches_message_plan(with_norm):
    torch.manual_seed(0)
    g = dgl.rand_graph(200, 3000)
    et = torch.randint(0, 5, (3000,))
    x = torch.randn(200, 16, requires_grad=True)
    norm = torch.rand(3000, 1) if with_norm else None
    conv = dgl.nn.RelGraphConv(16, 24, 5, regularizer="basis", num_bases=2)
    y = conv(g, x, et, norm)
    y.sum().backward()
    gx, gc, gw = x.grad.clone(), conv.linear_r.coeff.grad.clone(), conv.linear_r.W.grad.clone()
    x.grad = None
    conv.zero_grad()
    ref = _rgcn_message_plan(conv, g, x, et, norm)
    ref.sum().backward()
    assert torch.allclose(y, ref, atol=1e-4)
    assert torch.allclose(gx, x.grad, atol=1e-4) and torch.allclose(gc, conv.linear_r.coeff.grad, atol=1e-3)
    assert torch.allclose(gw, conv.linear_r.W.grad, atol=1e-3)


def test_dst_sorted_view_structure():
    g = dgl.rand_graph(50, 400)
    view, emap = g.dst_sorted_view()
    u, v = g.edges()
    vu, vv = view.edges()
    assert torch.equal(vu, u[emap.long()]) and torch.equal(vv, v[emap.long()])
    assert bool((vv[1:] >= vv[:-1]).all())
    assert view.dst_sorted_view()[1] is None          # already ordered: the view of a view is itself
    assert g.dst_sorted_view()[0]._graphs[0] is view._graphs[0]   # cached


@pytest.mark.gpu
def test_gatconv_cuda_view_matches_cpu():
    torch.manual_seed(0)
    g = dgl.add_self_loop(dgl.rand_graph(300, 5000))
    x = torch.randn(300, 32)
    conv = dgl.nn.GATConv(32, 8, 4)
    y, a = conv(g, x, get_attention=True)
    gc, xc = g.to("cuda"), x.cuda().requires_grad_()
    convc = conv.to("cuda")
    yc, ac = convc(gc, xc, get_attention=True)
    assert torch.allclose(yc.cpu(), y, atol=1e-4) and torch.allclose(ac.cpu(), a, atol=1e-5)
    yc.sum().backward()
    x2 = x.clone().requires_grad_()
    conv.cpu()(g, x2).sum().backward()
    assert torch.allclose(xc.grad.cpu(), x2.grad, atol=1e-4)


@pytest.mark.gpu
def test_sddmm_dot_edge_kernels_match_dense():
    g = dgl.rand_graph(400, 6000).to("cuda")
    u, v = g.edges()
    for shape in [(8, 8), (4, 16), (1, 41), (2, 64), (16, 4)]:
        a = torch.randn(400, *shape, device="cuda")
        b = torch.randn(400, *shape, device="cuda")
        got = dgl.ops.u_dot_v(g, a, b)
        ref = (a[u.long()] * b[v.long()]).sum(-1, keepdim=True)
        assert torch.allclose(got, ref, atol=1e-4), shape
    a = torch.randn(400, 8, 8, device="cuda").bfloat16()
    b = torch.randn(400, 8, 8, device="cuda").bfloat16()
    got = dgl.ops.u_dot_v(g, a, b).float()
    ref = (a[u.long()].float() * b[v.long()].float()).sum(-1, keepdim=True)
    assert torch.allclose(got, ref, atol=5e-2)


def test_trace_records_native_ops():
    g = dgl.rand_graph(100, 1000)
    x = torch.randn(100, 16, requires_grad=True)
    with dgl.utils.trace.trace() as t:
        with dgl.utils.trace.range("layer"):
            y = dgl.ops.copy_u_sum(g, x)
        y.sum().backward()
    s = t.summary()
    assert s["dglb200_cpu_gspmm"]["calls"] == 2 and s["range:layer"]["calls"] == 1
    assert not dgl.utils.trace.enabled() and "dglb200_cpu_gspmm" in t.report()


@pytest.mark.gpu
@pytest.mark.parametrize("H,D,dtype", [(8, 8, torch.float32), (1, 41, torch.float32), (4, 16, torch.bfloat16), (2, 32, torch.float32),
                                       (4, 64, torch.float32), (8, 64, torch.bfloat16), (3, 128, torch.float32)])   # column tiles
def test_fused_gat_matches_unfused(H, D, dtype):
    from dgl_b200.ops import gat_fused
    torch.manual_seed(0)
    g = dgl.add_self_loop(dgl.rand_graph(500, 9000)).to("cuda")
    view, _ = g.dst_sorted_view()
    ug = view._graphs[0]
    vec = 16 // torch.empty(0, dtype=dtype).element_size()
    Dp = D + (-D) % vec
    el = torch.randn(500, H, device="cuda").to(dtype).requires_grad_()
    er = torch.randn(500, H, device="cuda").to(dtype).requires_grad_()
    ft = torch.randn(500, H, Dp, device="cuda").to(dtype).requires_grad_()
    assert gat_fused.supported(el, ft, H, Dp)
    out = gat_fused.fused_gat(ug, el, er, ft, 0.2)
    go = torch.randn_like(out)
    out.backward(go)
    # oracle: dense fp32 formulation on the same edges
    u, v = view.edges()
    el32, er32, ft32 = (t.detach().float().requires_grad_() for t in (el, er, ft))
    e = torch.nn.functional.leaky_relu(el32[u.long()] + er32[v.long()], 0.2)              # [E, H]
    a = dgl.ops.edge_softmax(view, e.unsqueeze(-1)).squeeze(-1)
    ref = torch.zeros(500, H, Dp, device="cuda").index_add_(0, v.long(), a.unsqueeze(-1) * ft32[u.long()])
    ref.backward(go.float())
    tol = 2e-4 if dtype == torch.float32 else 6e-2
    assert torch.allclose(out.float(), ref, atol=tol)
    assert torch.allclose(ft.grad.float(), ft32.grad, atol=tol * 5)
    assert torch.allclose(el.grad.float(), el32.grad, atol=tol * 20) and torch.allclose(er.grad.float(), er32.grad, atol=tol * 20)


def test_graphbolt_hetero_labor_sampling():
    import dgl_b200.graphbolt as gb
    torch.manual_seed(0)
    u1, i1 = torch.randint(0, 50, (600,)), torch.randint(0, 30, (600,))
    i2, u2 = torch.randint(0, 30, (600,)), torch.randint(0, 50, (600,))
    hg = dgl.heterograph({("u", "r1", "i"): (u1, i1), ("i", "r2", "u"): (i2, u2)})
    fg = gb.from_dglgraph(hg)
    seeds = {"i": torch.arange(10), "u": torch.arange(5)}
    sub = fg.sample_layer_neighbors(seeds, [3], random_seed=5)
    sub2 = fg.sample_layer_neighbors(seeds, [3], random_seed=5)
    e1 = set(zip(u1.tolist(), i1.tolist()))
    csc = sub.sampled_csc["u:r1:i"]
    assert torch.equal(csc.indices, sub2.sampled_csc["u:r1:i"].indices)          # same seed -> same sample
    assert csc.indptr.shape[0] == 11
    for d in range(10):
        for s in csc.indices[int(csc.indptr[d]): int(csc.indptr[d + 1])].tolist():
            assert (s, d) in e1
    full = fg.sample_layer_neighbors(seeds, [-1])
    assert int(full.sampled_csc["u:r1:i"].indptr[-1]) == int(sum((i1 == d).sum() for d in range(10)))


@pytest.mark.gpu
def test_sparse_ops_under_autocast():
    g = dgl.add_self_loop(dgl.rand_graph(300, 4000)).to("cuda")
    x = torch.randn(300, 32, device="cuda")
    w = torch.rand(g.num_edges(), 1, device="cuda")                       # fp32 edge weights meet bf16 activations
    lin = torch.nn.Linear(32, 64).cuda()
    conv = dgl.nn.GATConv(64, 8, 4).cuda()
    rel = dgl.nn.RelGraphConv(32, 16, 3, regularizer="basis", num_bases=2).cuda()
    et = torch.randint(0, 3, (g.num_edges(),), device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        h = lin(x)
        assert h.dtype == torch.bfloat16
        agg = dgl.ops.u_mul_e_sum(g, h, w)
        assert agg.dtype == torch.bfloat16
        out = conv(g, h)
        out2 = rel(g, x, et)
        tl = dgl.nn.TypedLinear(32, 16, 3).cuda()
        y = tl(x, torch.sort(torch.randint(0, 3, (300,), device="cuda"))[0], sorted_by_type=True)
        assert y.dtype == torch.bfloat16                                  # tcgen05 grouped GEMM under autocast
    (agg.float().sum() + out.float().sum() + out2.float().sum() + y.float().sum()).backward()
    ref = dgl.ops.u_mul_e_sum(g, lin(x), w)
    assert (agg.float() - ref).abs().max() < 0.15 and lin.weight.grad is not None


@pytest.mark.gpu
def test_sageconv_block_uses_tcgen05_path_and_matches_unfused():
    from dgl_b200.dataloading import NeighborSampler
    from dgl_b200.nn import fused_sage
    torch.manual_seed(0)
    g = dgl.rand_graph(5000, 100000, idtype=torch.int32).to("cuda")
    feat = torch.randn(5000, 100, device="cuda")
    inp, out, blocks = NeighborSampler([10]).sample(g, torch.arange(700, device="cuda", dtype=torch.int32))
    blk = blocks[0]
    conv = dgl.nn.SAGEConv(100, 48, "mean").cuda()      # linear layer: gradients comparable without ReLU-mask flips
    x = feat[inp.long()].requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert fused_sage.usable(blk, x, 48, None)
        y = conv(blk, x)
        conv_r = dgl.nn.SAGEConv(100, 48, "mean", activation=torch.nn.functional.relu).cuda()
        conv_r.load_state_dict(conv.state_dict())
        yr = conv_r(blk, x.detach())                     # fused ReLU epilogue
    assert y.dtype == torch.bfloat16 and y.shape == (700, 48)
    assert (yr.float() - torch.relu(y.float())).abs().max() < 0.02
    gy = torch.randn_like(y)
    y.backward(gy)
    gx, gws, gwn, gb = x.grad.clone(), conv.fc_self.weight.grad.clone(), conv.fc_neigh.weight.grad.clone(), conv.fc_self.bias.grad.clone()
    # fp32 unfused oracle (plain formulation of the layer)
    x2 = x.detach().clone().requires_grad_()
    conv.zero_grad()
    hn = dgl.ops.copy_u_mean(blk, x2)
    ref = x2[:700] @ conv.fc_self.weight.t() + conv.fc_self.bias + hn @ conv.fc_neigh.weight.t()
    ref.backward(gy.float())
    assert (y.float() - ref).abs().max() < 0.08
    assert (gx - x2.grad).abs().max() < 0.08
    assert (gws - conv.fc_self.weight.grad).abs().max() / conv.fc_self.weight.grad.abs().max() < 0.03
    assert (gwn - conv.fc_neigh.weight.grad).abs().max() / conv.fc_neigh.weight.grad.abs().max() < 0.03
    assert (gb - conv.fc_self.bias.grad).abs().max() / conv.fc_self.bias.grad.abs().max() < 0.03


@pytest.mark.gpu
def test_fused_gat_zero_in_degree_rows():
    torch.manual_seed(0)
    g = dgl.rand_graph(3000, 6000, idtype=torch.int32)                    # avg in-degree 2: ~13 % of the rows are empty
    x = torch.randn(3000, 24)
    conv = dgl.nn.GATConv(24, 8, 4, allow_zero_in_degree=True)
    y = conv(g, x)
    zero = (g.in_degrees() == 0)
    assert int(zero.sum()) > 50
    yc = conv.cuda()(g.to("cuda"), x.cuda())
    assert bool(torch.isfinite(yc).all()) and torch.allclose(yc.cpu(), y, atol=1e-4)
    bias = conv.bias.detach().cpu().view(4, 8)
    assert torch.allclose(yc.cpu()[zero], bias.expand(int(zero.sum()), 4, 8), atol=1e-5)   # nothing aggregated: bias only


@pytest.mark.gpu
def test_deterministic_spmm_mode_is_bit_reproducible():
    from dgl_b200 import global_config as gc
    g = dgl.data.rmat_graph(20000, 600000, seed=1).to("cuda")                # power-law: hub rows far above the split degree
    x = torch.randn(20000, 64, device="cuda")
    assert int(g.in_degrees().max()) > 128
    ref = dgl.ops.copy_u_sum(g, x)
    gc.set_deterministic(True)
    try:
        a = dgl.ops.copy_u_sum(g, x)
        b = dgl.ops.copy_u_sum(g, x)
    finally:
        gc.set_deterministic(False)
    assert torch.equal(a, b) and torch.allclose(a, ref, atol=2e-2, rtol=1e-3)
