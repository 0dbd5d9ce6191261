"""Full-graph training benchmarks for BASELINE configs 3 and 4, runnable on BOTH arms with the same script:

  python benchmarks/bench_configs.py --config gat_reddit --impl ours       # dgl_b200
  python benchmarks/bench_configs.py --config gat_reddit --impl reference  # unmodified DGL from baseline/_ref

  gat_reddit : 2-layer GAT, 8 heads x 8 hidden, Reddit-shaped random graph (232,965 nodes / 114.6 M edges / 602 feats /
               41 classes) - the g-SDDMM + edge-softmax + u_mul_e g-SpMM path.
  rgcn_mag   : 2-layer R-GCN (RelGraphConv, basis regulariser, 2 bases, hidden 64) on the homogenised ogbn-mag-shaped
               heterograph (1.94 M nodes / 42.2 M typed edges incl. reverse relations, 8 relations, 349 classes).

The graph (torch CPU generator, fixed seed), the model definition (the `dgl.nn` API is the same on both arms), the
optimizer and the fp32 dtype are identical; a step = forward + loss + backward + Adam. Times are CUDA-event times after
3 warm-up epochs. One JSON line per run."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn as nn
import torch.nn.functional as F

MAG_NODES = {"paper": 736389, "author": 1134649, "institution": 8740, "field_of_study": 59965}
MAG_RELS = [("author", "affiliated_with", "institution", 1043998), ("author", "writes", "paper", 7145660),
            ("paper", "cites", "paper", 5416271), ("paper", "has_topic", "field_of_study", 7505078)]


def load_impl(impl):
    if impl == "ours":
        sys.path.insert(0, ROOT)
        import dgl_b200 as dgl
        return dgl
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    os.environ.setdefault("DGLBACKEND", "pytorch")
    import dgl
    import dgl.nn  # noqa: F401
    return dgl


def gat_reddit(dgl, scale, dev, dtype):
    n, e, f, c = int(232965 * scale), int(114615892 * scale), 602, 41
    gen = torch.Generator().manual_seed(0)
    u = torch.randint(0, n, (e,), generator=gen, dtype=torch.int32)
    v = torch.randint(0, n, (e,), generator=gen, dtype=torch.int32)
    g = dgl.graph((u, v), num_nodes=n, idtype=torch.int32)
    g = dgl.add_self_loop(g).to(dev)
    feat = torch.randn(n, f, generator=gen).to(dev).to(dtype)
    labels = torch.randint(0, c, (n,), generator=gen).to(dev)
    train = torch.randperm(n, generator=gen)[: int(0.66 * n)].to(dev)

    class GAT(nn.Module):
        def __init__(self):
            super().__init__()
            self.l1 = dgl.nn.GATConv(f, 8, 8, activation=F.elu)
            self.l2 = dgl.nn.GATConv(64, c, 1)

        def forward(self, g, x):
            h = self.l1(g, x).flatten(1)
            return self.l2(g, h).mean(1)

    model = GAT().to(dev).to(dtype)
    return model, (lambda: model(g, feat)), labels, train, dict(nodes=n, edges=g.num_edges(), feats=f, classes=c, heads=8, hidden=8)


def gcn_cora(dgl, scale, dev, dtype):
    """BASELINE config 1: 2-layer GCN (16 hidden) on a Cora-sized random graph (2,708 nodes / 10,556 edges / 1,433 feats /
    7 classes) - the CPU plumbing configuration."""
    n, e, f, c = int(2708 * scale), int(10556 * scale), 1433, 7
    gen = torch.Generator().manual_seed(0)
    u = torch.randint(0, n, (e,), generator=gen, dtype=torch.int32)
    v = torch.randint(0, n, (e,), generator=gen, dtype=torch.int32)
    g = dgl.add_self_loop(dgl.graph((u, v), num_nodes=n, idtype=torch.int32)).to(dev)
    feat = (torch.rand(n, f, generator=gen) < 0.0127).float().to(dev).to(dtype)      # Cora's bag-of-words sparsity
    labels = torch.randint(0, c, (n,), generator=gen).to(dev)
    train = torch.arange(140).to(dev)

    class GCN(nn.Module):
        def __init__(self):
            super().__init__()
            self.l1 = dgl.nn.GraphConv(f, 16, activation=F.relu)
            self.l2 = dgl.nn.GraphConv(16, c)

        def forward(self, g, x):
            return self.l2(g, F.dropout(self.l1(g, x), 0.5, self.training))

    model = GCN().to(dev).to(dtype)
    return model, (lambda: model(g, feat)), labels, train, dict(nodes=n, edges=g.num_edges(), feats=f, classes=c, hidden=16)


def rgcn_mag(dgl, scale, dev, dtype):
    gen = torch.Generator().manual_seed(0)
    nn_ = {k: max(8, int(v * scale)) for k, v in MAG_NODES.items()}
    data = {}
    for s, r, d, e in MAG_RELS:
        e = max(16, int(e * scale))
        u = torch.randint(0, nn_[s], (e,), generator=gen, dtype=torch.int32)
        v = torch.randint(0, nn_[d], (e,), generator=gen, dtype=torch.int32)
        data[(s, r, d)] = (u, v)
        data[(d, "rev_" + r, s)] = (v, u)
    hg = dgl.heterograph(data, num_nodes_dict=nn_, idtype=torch.int32)
    g = dgl.to_homogeneous(hg)
    # edges sorted by relation: both arms then take their segment-MM fast path (presorted=True)
    et = g.edata[dgl.ETYPE]
    order = torch.sort(et.long(), stable=True)[1]
    uu, vv = g.edges()
    g2 = dgl.graph((uu[order], vv[order]), num_nodes=g.num_nodes(), idtype=torch.int32)
    ntypes = g.ndata[dgl.NTYPE]
    g2 = g2.to(dev)
    etypes = et[order].to(dev)
    num_rels = len(hg.canonical_etypes)
    C, H = 349, 64
    target = (ntypes == hg.ntypes.index("paper")).nonzero().squeeze(1).to(dev)
    labels = torch.randint(0, C, (target.numel(),), generator=gen).to(dev)
    N = g2.num_nodes()

    class RGCN(nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = nn.Embedding(N, H)
            self.l1 = dgl.nn.RelGraphConv(H, H, num_rels, regularizer="basis", num_bases=2, self_loop=True)
            self.l2 = dgl.nn.RelGraphConv(H, C, num_rels, regularizer="basis", num_bases=2, self_loop=True)

        def forward(self):
            h = F.relu(self.l1(g2, self.emb.weight, etypes, presorted=True))
            return self.l2(g2, h, etypes, presorted=True)

    model = RGCN().to(dev).to(dtype)
    return model, (lambda: model()), labels, target, dict(nodes=N, edges=g2.num_edges(), relations=num_rels, hidden=H, classes=C, bases=2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="gat_reddit", choices=["gat_reddit", "rgcn_mag", "gcn_cora"])
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device(args.device)
    dgl = load_impl(args.impl)
    dtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    model, fwd, labels, idx, cfg = {"gat_reddit": gat_reddit, "rgcn_mag": rgcn_mag, "gcn_cora": gcn_cora}[args.config](dgl, args.scale, dev, dtype)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)

    def step():
        logits = fwd()
        loss = F.cross_entropy(logits[idx].float(), labels[idx] if labels.shape[0] == logits.shape[0] else labels)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        loss = step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.epochs):
            loss = step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / args.epochs
    else:
        import time
        t0 = time.perf_counter()
        for _ in range(args.epochs):
            loss = step()
        ms = (time.perf_counter() - t0) * 1e3 / args.epochs
    rec = dict(bench=args.config, impl=args.impl, dtype=args.dtype, scale=args.scale, ms_per_epoch=ms,
               edges_per_s=cfg["edges"] / ms * 1e3, final_loss=float(loss), peak_mem_gb=(torch.cuda.max_memory_allocated() / 2 ** 30) if dev.type == "cuda" else None, device=dev.type,
               config=cfg, epochs=args.epochs, warmup=args.warmup)
    print(json.dumps(rec), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
