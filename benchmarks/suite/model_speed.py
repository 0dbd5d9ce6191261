"""Model-speed benchmarks (reference benchmarks/model_speed/: bench_gcn, bench_gat, bench_sage, bench_sage_ns, bench_rgcn_*)."""
import torch
import torch.nn.functional as F

import dgl_b200 as dgl
from dgl_b200.models import GAT, GCN, RGCN, GraphSAGE

from . import bench, timeit


def _train_step(model, forward, labels, idx=None):
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)

    def step():
        out = forward()
        loss = F.cross_entropy(out if idx is None else out[idx], labels if idx is None else labels[idx])
        opt.zero_grad()
        loss.backward()
        opt.step()
    return step


@bench("model_speed", model=["gcn", "gat", "sage"])
def fullgraph_epoch(device, model):
    g = dgl.add_self_loop(dgl.data.rmat_graph(50_000, 500_000)).to(device)
    x = torch.randn(g.num_nodes(), 128, device=device)
    y = torch.randint(0, 10, (g.num_nodes(),), device=device)
    if model == "gcn":
        m = GCN(128, 64, 10).to(device)
    elif model == "gat":
        m = GAT(128, 8, 10, heads=(8, 1)).to(device)
    else:
        m = GraphSAGE(128, 64, 10, 2, 0.5).to(device)
        return timeit(_train_step(m, lambda: m([g, g], x), y), device)
    return timeit(_train_step(m, lambda: m(g, x), y), device)


@bench("model_speed")
def rgcn_epoch(device):
    g = dgl.data.rmat_graph(30_000, 300_000).to(device)
    et = torch.randint(0, 8, (g.num_edges(),), device=device)
    y = torch.randint(0, 5, (g.num_nodes(),), device=device)
    m = RGCN(g.num_nodes(), 32, 5, 8, num_bases=2).to(device)
    return timeit(_train_step(m, lambda: m(g, et), y), device)


@bench("model_speed", fanout=[[10, 10], [15, 10, 5]])
def sage_neighbor_sampling_step(device, fanout):
    from dgl_b200.dataloading import DataLoader, NeighborSampler
    g = dgl.data.rmat_graph(100_000, 2_000_000).to(device)
    x = torch.randn(g.num_nodes(), 100, device=device)
    y = torch.randint(0, 47, (g.num_nodes(),), device=device)
    m = GraphSAGE(100, 256, 47, len(fanout), 0.5).to(device)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    dl = DataLoader(g, torch.arange(50_000, device=device, dtype=g.idtype), NeighborSampler(fanout), batch_size=1024,
                    shuffle=True)
    it = iter(dl)

    def step():
        nonlocal it
        try:
            inp, out, blocks = next(it)
        except StopIteration:
            it = iter(dl)
            inp, out, blocks = next(it)
        loss = F.cross_entropy(m(blocks, x[inp.long()]), y[out.long()])
        opt.zero_grad()
        loss.backward()
        opt.step()
    return timeit(step, device, warmup=2, iters=10)


@bench("model_speed")
def sage_fused_engine_step(device):
    if not str(device).startswith("cuda"):
        return float("nan")
    from dgl_b200.models import FusedSAGETrainer
    g = dgl.data.rmat_graph(100_000, 2_000_000).to(device)
    x = torch.randn(g.num_nodes(), 100, device=device)
    y = torch.randint(0, 47, (g.num_nodes(),), device=device)
    tr = FusedSAGETrainer(g, x, y, 47, hidden=256, fanouts=[10, 10, 10], batch_size=1024)
    seeds = [torch.randint(0, 100_000, (1024,)) for _ in range(16)]
    k = [0]

    def step():
        tr.step_async(seeds[k[0] % 16])
        k[0] += 1
    return timeit(step, device, warmup=5, iters=50)
