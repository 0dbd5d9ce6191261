"""GraphBolt GPU path, GPU cache kernel, P2P comm kernels in single-rank mode."""
import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.graphbolt as gb

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_graphbolt_gpu_pipeline():
    g = dgl.rand_graph(3000, 60000)
    fg = gb.from_dglgraph(g).to(DEV)
    feat = torch.randn(3000, 32)
    store = gb.TorchBasedFeatureStore({("node", None, "feat"): feat}).pin_memory_()
    items = gb.ItemSet((torch.arange(500), torch.randint(0, 3, (500,))), names=("seeds", "labels"))
    dp = gb.ItemSampler(items, batch_size=128, shuffle=True).copy_to(DEV).sample_neighbor(fg, [4, 4]).fetch_feature(store, ["feat"])
    n = 0
    for mb in gb.DataLoader(dp, num_workers=1):
        assert mb.node_features["feat"].is_cuda
        assert torch.equal(mb.node_features["feat"].cpu(), feat[mb.input_nodes.cpu()])
        b = mb.blocks
        h = dgl.nn.SAGEConv(32, 8, "mean").to(DEV)(b[0], mb.node_features["feat"])
        assert h.shape[0] == b[0].num_dst_nodes()
        sg = mb.sampled_subgraphs[-1]
        src = sg.original_row_node_ids[sg.sampled_csc.indices.long()]
        deg = sg.sampled_csc.indptr[1:] - sg.sampled_csc.indptr[:-1]
        dst = torch.repeat_interleave(sg.original_column_node_ids, deg.long())
        assert bool(g.has_edges_between(src.cpu(), dst.cpu()).all())
        n += 1
    assert n == 4


def test_gpu_cache_kernel():
    base = torch.arange(5000, dtype=torch.float32).unsqueeze(1).repeat(1, 16)
    fb = gb.TorchBasedFeature(base.clone()).pin_memory_()
    cf = gb.GPUCachedFeature(fb, 512 * 16 * 4)
    for it in range(20):
        k = (torch.randn(600).abs() * 300).long().clamp(max=4999).to(DEV)
        out = cf.read(k)
        assert torch.equal(out.cpu(), base[k.cpu()])
    assert cf.miss_rate < 0.9
    c = gb.GPUCache((64, 4), torch.float32)
    keys = torch.arange(40, device=DEV)
    c.replace(keys, torch.ones(40, 4, device=DEV) * keys.unsqueeze(1))
    out, miss, mk = c.query(torch.tensor([3, 7, 1000], device=DEV))
    assert miss.tolist() == [2] and out[0, 0] == 3 and out[1, 0] == 7


def test_p2p_single_rank_kernels():
    from dgl_b200.parallel import P2PComm
    comm = P2PComm()
    n, F = 2000, 64
    shard = comm.alloc(n * F * 2)
    x = shard.view(torch.bfloat16)[: n * F].view(n, F)
    x.copy_(torch.randn(n, F, device=DEV))
    g = dgl.rand_graph(n, 30000, idtype=torch.int32).to(DEV)
    indptr, indices, _ = g._graphs[0].csc()
    out = torch.empty(n, F, dtype=torch.bfloat16, device=DEV)
    comm.halo_spmm(indptr, indices, shard, n, F, torch.bfloat16, out, mean=True)
    ref = dgl.ops.copy_u_mean(g, x.float())
    assert (out.float() - ref).abs().max() < 0.05
    plan = comm.halo_plan(indptr, indices, n)      # single rank: no halo, every row is interior
    pbuf, table = plan.alloc_features(F, torch.bfloat16)
    table[:n].copy_(x)
    out2 = torch.empty_like(out)
    plan.spmm(pbuf, table, out2, mean=True)
    assert plan.n_halo == 0 and (out2.float() - ref).abs().max() < 0.05
    dsh = comm.alloc(n * F * 4)
    dx = dsh.view(torch.float32)[: n * F].view(n, F)
    dout = torch.randn(n, F, device=DEV)
    comm.halo_spmm_backward(indptr, indices, dsh, n, F, dout, mean=False)
    u, v = g.edges()
    ref = torch.zeros(n, F, device=DEV).index_add_(0, u.long(), dout[v.long()])
    assert torch.allclose(dx, ref, atol=1e-3)
    buf = comm.alloc(1024 * 4)
    buf.view(torch.float32)[:1024].copy_(torch.arange(1024.0))
    res = torch.empty(1024, device=DEV)
    comm.all_reduce(buf, res, 1024, scale=0.5)
    assert torch.allclose(res, torch.arange(1024.0, device=DEV) * 0.5)
    idx = torch.randint(0, n, (500,), device=DEV)
    got = torch.empty(500, F, dtype=torch.bfloat16, device=DEV)
    comm.gather(shard, n, idx, got)
    assert torch.equal(got, x[idx])
    comm.check_error()


def test_native_gpu_graph_cache_and_index_select_csc():
    import dgl_b200.graphbolt as gb
    from dgl_b200.graphbolt.impl import GPUGraphCache, index_select_csc
    g = dgl.rand_graph(2000, 40000, idtype=torch.int64)
    indptr, indices, _ = g._graphs[0].csc()
    indptr_p, indices_p = indptr.pin_memory(), indices.pin_memory()          # graph stays in host memory (UVA)
    seeds = torch.randint(0, 2000, (500,), device=DEV)
    ip, ix = index_select_csc(indptr_p, indices_p, seeds)
    ip_ref, ix_ref = index_select_csc(indptr, indices, seeds.cpu())
    assert torch.equal(ip.cpu(), ip_ref) and torch.equal(ix.cpu(), ix_ref)
    cache = GPUGraphCache(20000, threshold=2, dtypes=[torch.int64])
    hot = torch.arange(0, 300, device=DEV)
    for rnd in range(3):
        hit, miss = cache.query(hot)
        if rnd == 0:
            assert hit.numel() == 0 and miss.numel() == 300
        ms = hot[miss]
        mp, mi = index_select_csc(indptr_p, indices_p, ms)
        cache.replace(ms, mp, mi)
    hit, miss = cache.query(hot)
    assert hit.numel() > 250                                                # admitted after reaching the threshold
    cp, ci = cache.fetch_cached(hot[hit])
    rp, ri = index_select_csc(indptr, indices, hot[hit].cpu())
    assert torch.equal(cp.cpu(), rp) and torch.equal(ci.cpu(), ri)
    assert 0.0 < cache.miss_rate < 1.0 and cache.used <= 20000


# ----------------------------------------------------------------------------------- whole-step CUDA graph, generic nn.Module
def test_graphed_trainer_matches_eager_semantics_and_learns():
    """GraphedTrainer: the loss is finite, replays draw different samples, partial batches ignore the padded slots and
    the model learns a planted-partition task."""
    import dgl_b200 as dgl
    from dgl_b200.dataloading import GraphedTrainer
    from dgl_b200.models import GraphSAGE
    torch.manual_seed(0)
    n, C = 3000, 6
    centres = torch.randn(C, 16) * 2
    y = torch.randint(0, C, (n,))
    x = centres[y] + torch.randn(n, 16)
    # homophilous graph: most edges inside a class
    src = torch.randint(0, n, (40000,))
    same = torch.rand(40000) < 0.8
    pool = [torch.nonzero(y == c).squeeze(1) for c in range(C)]
    dst = torch.where(same, torch.stack([pool[int(c)][torch.randint(0, len(pool[int(c)]), (1,))][0] for c in y[src]]),
                      torch.randint(0, n, (40000,)))
    g = dgl.graph((src.int(), dst.int()), num_nodes=n, idtype=torch.int32).to("cuda")
    g.create_formats_()
    model = GraphSAGE(16, 64, C, 2, 0.0).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, capturable=True)
    tr = GraphedTrainer(g, x.cuda(), y.cuda(), model, opt, [5, 5], 256, amp=False)
    seeds = torch.randperm(n)[:256]
    first = float(tr.step(seeds))
    assert first == first and first > 0
    e1 = tr.sampled_edges()
    # (2) same seeds, next replay: other neighbours (the device step counter keys the RNG)
    nbr_a = tr.ws.nbr_g[2].clone()
    tr.step(seeds)
    assert not torch.equal(nbr_a, tr.ws.nbr_g[2]) and tr.sampled_edges() > e1
    # (3) a partial batch: padded slots do not contribute
    part = torch.randperm(n)[:100]
    lp = float(tr.step(part))
    assert lp == lp and int(tr.ws.n[2]) == 100
    # (4) training reduces the loss and reaches a sensible accuracy
    losses = []
    for ep in range(6):
        perm = torch.randperm(n)
        for i in range(0, n - 255, 256):
            losses.append(tr.step(perm[i:i + 256]).clone())
    losses = torch.stack(losses).cpu()
    assert losses[-5:].mean() < 0.6 * losses[:5].mean(), (losses[:5], losses[-5:])
    model.eval()
    with torch.no_grad():
        pred = model.full_forward(g, x.cuda()).argmax(1).cpu() if hasattr(model, "full_forward") else None
    if pred is not None:
        assert (pred == y).float().mean() > 0.8


def test_graphed_trainer_table_input_mode_learns():
    """gather_inputs=False: the first SAGEConv reads the feature table in place through global ids."""
    import dgl_b200 as dgl
    from dgl_b200.dataloading import GraphedTrainer
    from dgl_b200.models import GraphSAGE
    torch.manual_seed(1)
    n, C, Fdim = 4000, 5, 64
    y = torch.randint(0, C, (n,))
    x = (torch.randn(C, Fdim) * 1.5)[y] + torch.randn(n, Fdim)
    src = torch.randint(0, n, (60000,))
    dst = torch.randint(0, n, (60000,))
    g = dgl.graph((src.int(), dst.int()), num_nodes=n, idtype=torch.int32).to("cuda")
    g.create_formats_()
    for amp in (False, True):
        model = GraphSAGE(Fdim, 64, C, 2, 0.0).cuda()
        opt = torch.optim.Adam(model.parameters(), lr=5e-3, capturable=True)
        tr = GraphedTrainer(g, x.cuda(), y.cuda(), model, opt, [4, 4], 512, amp=amp, gather_inputs=False)
        losses = []
        for ep in range(5):
            perm = torch.randperm(n)
            for i in range(0, n - 511, 512):
                losses.append(tr.step(perm[i:i + 512]).clone())
        losses = torch.stack(losses).cpu()
        assert torch.isfinite(losses).all() and losses[-4:].mean() < 0.5 * losses[:4].mean(), (amp, losses[:4], losses[-4:])
    with pytest.raises(dgl.DGLError):
        GraphedTrainer(g, x[:, :20].cuda(), y.cuda(), model, opt, [4, 4], 512, gather_inputs=False)


def _capacity_and_real_blocks(n_src=600, n_dst=200, fanout=6, seed=0):
    """One sampled layer as (CapacityBlock with padding rows, equivalent DGLBlock)."""
    from dgl_b200.dataloading.graphed import CapacityBlock
    g = torch.Generator().manual_seed(seed)
    cap = n_dst + 56                                             # padding rows beyond the live count
    cnt = torch.randint(0, fanout + 1, (cap,), generator=g).int()
    nbr = torch.randint(0, n_src, (cap, fanout), generator=g).int()
    nbr[torch.arange(fanout).unsqueeze(0) >= cnt.unsqueeze(1)] = -1
    n_dev = torch.tensor([n_dst], dtype=torch.int64)
    cb = CapacityBlock(nbr.cuda(), cnt.cuda(), fanout, n_src, cap, n_dev.cuda())
    rows = torch.arange(n_dst).repeat_interleave(cnt[:n_dst].long())
    cols = nbr[:n_dst][nbr[:n_dst] >= 0].long()
    block = dgl.create_block((cols, rows), num_src_nodes=n_src, num_dst_nodes=n_dst).to("cuda")
    return cb, block, n_dst


@pytest.mark.parametrize("layer", ["gcn_both", "gcn_right", "gcn_none", "gcn_left", "gin_sum", "gin_mean", "sage_wide", "sage_gcn",
                                   "gat", "gat_residual"])
def test_capacity_block_layers_match_regular_blocks(layer):
    """GraphConv / GINConv / SAGEConv on a capacity block (ELL kernels, device-side degrees) equal the same layer on the
    equivalent variable-shape block: outputs of the live rows and all gradients."""
    import copy
    import dgl_b200.nn as dglnn
    torch.manual_seed(0)
    cb, block, n_dst = _capacity_and_real_blocks()
    F_in, F_out = 20, 12
    if layer.startswith("gcn"):
        m = dglnn.GraphConv(F_in, F_out, norm=layer.split("_")[1], allow_zero_in_degree=True)
    elif layer.startswith("gin"):
        m = dglnn.GINConv(torch.nn.Linear(F_in, F_out), layer.split("_")[1], learn_eps=True)
    elif layer.startswith("gat"):
        m = dglnn.GATConv(F_in, F_out // 4, 4, residual=layer.endswith("residual"), allow_zero_in_degree=True)
    elif layer == "sage_wide":
        F_in = 320                                             # wider than the tcgen05 layer takes: ELL fallback
        m = dglnn.SAGEConv(F_in, F_out, "mean")
    else:
        m = dglnn.SAGEConv(F_in, F_out, "gcn")
    m = m.cuda()
    ref = copy.deepcopy(m)
    x = torch.randn(600, F_in, device="cuda")
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    out_c = m(cb, xa)[:n_dst]
    out_r = ref(block, xb)
    if out_r.dim() == 3:
        out_c, out_r = out_c.flatten(1), out_r.flatten(1)
    assert out_c.shape == out_r.shape and torch.allclose(out_c, out_r, atol=2e-4), (out_c - out_r).abs().max()
    go = torch.randn_like(out_r)
    out_c.backward(go)
    out_r.backward(go)
    assert torch.allclose(xa.grad, xb.grad, atol=2e-4)
    for pa, pb in zip(m.parameters(), ref.parameters()):
        assert torch.allclose(pa.grad, pb.grad, atol=5e-4), (pa.grad - pb.grad).abs().max()


def test_graphed_trainer_with_gcn_and_gin_models():
    """The whole-step CUDA graph is not tied to SAGE: a GCN stack and a GIN stack built from the public nn layers train."""
    import dgl_b200.nn as dglnn
    from dgl_b200.dataloading import GraphedTrainer
    torch.manual_seed(0)
    n, C, Fdim = 3000, 4, 32
    y = torch.arange(n) % C                                        # class c = ids congruent to c: easy same-class draws
    x = (torch.randn(C, Fdim) * 1.5)[y] + torch.randn(n, Fdim)
    dst = torch.randint(0, n, (50000,))
    same = torch.rand(50000) < 0.85                                # homophilous: neighbours mostly share the label
    src = torch.where(same, torch.randint(0, n // C, (50000,)) * C + y[dst], torch.randint(0, n, (50000,)))
    g = dgl.graph((src.int(), dst.int()), num_nodes=n, idtype=torch.int32).to("cuda")
    g.create_formats_()

    class Net(torch.nn.Module):
        def __init__(self, kind):
            super().__init__()
            if kind == "gcn":
                self.l1 = dglnn.GraphConv(Fdim, 48, norm="both", allow_zero_in_degree=True)
                self.l2 = dglnn.GraphConv(48, 48, norm="right", allow_zero_in_degree=True)
            else:
                self.l1 = dglnn.GINConv(torch.nn.Linear(Fdim, 48), "mean")
                self.l2 = dglnn.GINConv(torch.nn.Linear(48, 48), "sum")
            self.out = torch.nn.Linear(48, C)

        def forward(self, blocks, h):
            h1 = torch.relu(self.l1(blocks[0], h))
            h2 = torch.relu(self.l2(blocks[1], h1))
            return self.out(h2)

    for kind in ("gcn", "gin"):
        model = Net(kind).cuda()
        opt = torch.optim.Adam(model.parameters(), lr=5e-3, capturable=True)
        tr = GraphedTrainer(g, x.cuda(), y.cuda(), model, opt, [5, 5], 256)
        losses = []
        for ep in range(6):
            perm = torch.randperm(n)
            for i in range(0, n - 255, 256):
                losses.append(tr.step(perm[i:i + 256]).clone())
        losses = torch.stack(losses).cpu()
        # (the capture warm-up already runs a few optimizer steps, so an easy task may start near its floor)
        assert torch.isfinite(losses).all() and losses[-5:].mean() < max(0.7 * losses[:5].mean(), 0.15), (kind, losses[:5], losses[-5:])
