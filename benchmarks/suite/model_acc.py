"""Model-accuracy benchmarks (reference benchmarks/model_acc/: bench_gcn, bench_gat, bench_sage, bench_sage_ns,
bench_rgcn_*). No network here, so the datasets are synthetic with a planted signal (labels are a function of the
aggregated features); the metric is the test accuracy after a fixed number of epochs, returned in place of seconds."""
import torch
import torch.nn.functional as F

import dgl_b200 as dgl
from dgl_b200.models import GAT, GCN, GraphSAGE

from . import bench


def _planted(device, n=5000, e=60000, f=32, c=5, seed=0):
    gen = torch.Generator().manual_seed(seed)
    g = dgl.add_self_loop(dgl.to_bidirected(dgl.rand_graph(n, e)))
    x = torch.randn(n, f, generator=gen)
    w = torch.randn(f, c, generator=gen)
    y = (dgl.ops.copy_u_mean(g, x) @ w + 0.5 * x @ w).argmax(1)      # label depends on the neighbourhood
    perm = torch.randperm(n, generator=gen)
    return g.to(device), x.to(device), y.to(device), perm[: n // 2].to(device), perm[n // 2:].to(device)


@bench("model_acc", model=["gcn", "gat", "sage"])
def planted_partition_accuracy(device, model):
    g, x, y, tr, te = _planted(device)
    torch.manual_seed(0)
    if model == "gcn":
        m = GCN(32, 64, 5).to(device)
        fwd = lambda: m(g, x)  # noqa: E731
    elif model == "gat":
        m = GAT(32, 8, 5, heads=(8, 1), feat_drop=0.0, attn_drop=0.0).to(device)
        fwd = lambda: m(g, x)  # noqa: E731
    else:
        m = GraphSAGE(32, 64, 5, 2, 0.0).to(device)
        fwd = lambda: m([g, g], x)  # noqa: E731
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    for _ in range(60):
        loss = F.cross_entropy(fwd()[tr], y[tr])
        opt.zero_grad()
        loss.backward()
        opt.step()
    m.eval()
    with torch.no_grad():
        return float((fwd()[te].argmax(1) == y[te]).float().mean())
