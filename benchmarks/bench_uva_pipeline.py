"""BASELINE config 5 (1 GPU slice): GraphSAGE mini-batch training where the graph AND the features stay in pinned host
memory; sampling and the feature gather run on the GPU reading host memory directly (UVA).

  --impl ours       GraphBolt pipeline of dgl_b200: ItemSampler -> sample_neighbor on a pinned FusedCSCSamplingGraph ->
                    fetch_feature from a pinned TorchBasedFeature (optionally behind the GPU cache) -> SAGE on blocks
  --impl reference  unmodified DGL: dgl.dataloading.DataLoader(use_uva=True) + NeighborSampler on the pinned graph,
                    features through dgl.utils.gather_pinned_tensor_rows, dgl.nn.SAGEConv model
                    (the reference's GraphBolt library is not built in this image; this is its UVA path in libdgl)

papers100M shape scaled by --scale (default 0.1: 11.1 M nodes, 161 M edges, 128 fp32 features = 5.7 GB pinned), skewed
random graph generated with torch (identical for both arms), batch 1024, fan-out 10,10,10, hidden 256, Adam.
Wall-clock per step is what matters here (the pipeline is host-driven); measured with CUDA events around K steps after W
warm-up steps, plus the device->host loss read of every step. One JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn as nn
import torch.nn.functional as F


_KEEP = []


def make_model(nnmod, kind, in_f=128, hid=256, out_f=172, heads=4):
    """The same 3-layer mini-batch model on either arm (``nnmod`` = dgl.nn or dgl_b200.nn): GraphSAGE-mean or GAT."""

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            if kind == "gat":
                d = hid // heads
                self.layers = nn.ModuleList([nnmod.GATConv(in_f, d, heads, allow_zero_in_degree=True),
                                             nnmod.GATConv(hid, d, heads, allow_zero_in_degree=True),
                                             nnmod.GATConv(hid, out_f, 1, allow_zero_in_degree=True)])
            else:
                self.layers = nn.ModuleList([nnmod.SAGEConv(in_f, hid, "mean"), nnmod.SAGEConv(hid, hid, "mean"),
                                             nnmod.SAGEConv(hid, out_f, "mean")])
            self.dropout = nn.Dropout(0.5)

        def forward(self, blocks, x):
            h = x
            for l, (layer, block) in enumerate(zip(self.layers, blocks)):
                h = layer(block, h)
                h = h.flatten(1) if h.dim() == 3 else h
                if l != 2:
                    h = self.dropout(F.relu(h))
            return h

    return Net()


def make_graph(scale):
    n = int(111_059_956 * scale)
    e = int(1_615_685_872 * scale)
    gen = torch.Generator().manual_seed(0)
    # skewed endpoints (square of a uniform) -> hubs, like a citation graph
    u = (torch.rand(e, generator=gen) ** 2 * n).long().clamp_(max=n - 1)
    v = torch.randint(0, n, (e,), generator=gen)
    feat = torch.randn(n, 128, generator=gen)
    labels = torch.randint(0, 172, (n,), generator=gen)
    train = torch.randperm(n, generator=gen)[: max(2048, int(0.011 * n))]
    return n, u, v, feat, labels, train


def run_ours(args, n, u, v, feat, labels, train):
    sys.path.insert(0, ROOT)
    import dgl_b200 as dgl
    import dgl_b200.graphbolt as gb
    from dgl_b200.models import GraphSAGE
    dev = "cuda"
    g = dgl.graph((u.int(), v.int()), num_nodes=n, idtype=torch.int32)
    if args.graphed:
        # the same nn.Module model and torch optimizer, but the whole step (sampling .. optimizer) replayed as ONE CUDA graph
        from dgl_b200.dataloading import GraphedTrainer
        g = g.to(dev)
        g.create_formats_()
        model = (GraphSAGE(128, 256, 172, 3, 0.5) if args.model == "sage" else make_model(dgl.nn, args.model)).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
        trainer = GraphedTrainer(g, feat.to(dev), labels.to(dev), model, opt, args.fanouts, args.batch, amp=args.amp,
                                 gather_inputs=not args.table_input)
        args._trainer = trainer

        def seed_batches():
            while True:
                perm = train[torch.randperm(train.numel())]
                for i in range(0, perm.numel(), args.batch):
                    yield perm[i:i + args.batch]
        return model, opt, seed_batches(), None, None
    if args.resident == "device":          # graph + features in HBM: the generic (non-fused) GPU pipeline
        fg = gb.from_dglgraph(g).to(dev)
        feature = gb.TorchBasedFeature(feat.to(dev))
    else:
        fg = gb.from_dglgraph(g).pin_memory_()
        feature = gb.TorchBasedFeature(feat).pin_memory_()
    if args.cache_mb and args.resident != "device":
        feature = gb.GPUCachedFeature(feature, args.cache_mb << 20)
    store = gb.BasicFeatureStore({("node", None, "feat"): feature})
    items = gb.ItemSet((train.int(), labels[train]), names=("seeds", "labels"))
    dp = gb.ItemSampler(items, batch_size=args.batch, shuffle=True).copy_to(dev)
    dp = dp.sample_neighbor(fg, args.fanouts).fetch_feature(store, ["feat"])
    # with the prefetch thread the UVA gather is capped to 10240 threads (gb default) so that it leaves SMs to the model
    loader = gb.DataLoader(dp, num_workers=args.workers, max_uva_threads=10240 if args.workers else 0)
    model = (GraphSAGE(128, 256, 172, 3, 0.5) if args.model == "sage" else make_model(dgl.nn, args.model)).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def batches():
        while True:
            for mb in loader:
                yield mb.blocks, mb.node_features["feat"], mb.labels.long().to(dev)

    return model, opt, batches(), lambda blocks: sum(b.num_edges() for b in blocks), lambda blocks: blocks[0].num_src_nodes()


def run_reference(args, n, u, v, feat, labels, train):
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    os.environ.setdefault("DGLBACKEND", "pytorch")
    import dgl
    import dgl.nn as dglnn
    dev = torch.device("cuda")
    g = dgl.graph((u.int(), v.int()), num_nodes=n, idtype=torch.int32)
    g.create_formats_()
    device_resident = args.resident == "device"
    if device_resident:
        g = g.to(dev)
        feat_d = feat.to(dev)
    else:
        g.pin_memory_()
        nfeat = dgl.utils.pin_memory_inplace(feat)
        _KEEP.append(nfeat)   # the holder un-pins the tensor when it is collected
    labels_d = labels.to(dev)
    sampler = dgl.dataloading.NeighborSampler(args.fanouts)
    dl = dgl.dataloading.DataLoader(g, train.int().to(dev), sampler, device=dev, batch_size=args.batch, shuffle=True,
                                    drop_last=False, num_workers=0, use_uva=not device_resident)

    model = make_model(dglnn, args.model).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def batches():
        while True:
            for input_nodes, output_nodes, blocks in dl:
                x = feat_d[input_nodes.long()] if device_resident else dgl.utils.gather_pinned_tensor_rows(feat, input_nodes)
                yield blocks, x, labels_d[output_nodes.long()]

    return model, opt, batches(), lambda blocks: sum(b.num_edges() for b in blocks), lambda blocks: blocks[0].num_src_nodes()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--cache-mb", type=int, default=0)
    ap.add_argument("--workers", type=int, default=1, help="ours: 1 = prefetch thread + side stream (gb.DataLoader)")
    ap.add_argument("--resident", default="host", choices=["host", "device"],
                    help="host: graph + features pinned in host memory (UVA, config 5); device: both in HBM")
    ap.add_argument("--amp", action="store_true", help="bf16 autocast (ours: SAGEConv takes the tcgen05 fused layer)")
    ap.add_argument("--table-input", action="store_true", help="with --graphed: the first SAGEConv reads the feature table in place")
    ap.add_argument("--graphed", action="store_true", help="ours, device resident: dataloading.GraphedTrainer (whole step as one CUDA graph)")
    ap.add_argument("--model", default="sage", choices=["sage", "gat"], help="3-layer GraphSAGE-mean (default) or 4-head GAT")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    args.fanouts = [10, 10, 10]
    t0 = time.time()
    n, u, v, feat, labels, train = make_graph(args.scale)
    gen_s = time.time() - t0
    model, opt, it, count_edges, count_inputs = (run_ours if args.impl == "ours" else run_reference)(args, n, u, v, feat, labels, train)

    def graphed_step():
        tr = args._trainer
        before = tr.edges.clone()
        loss = tr.step(next(it))
        return float(loss), int((tr.edges - before).item()), 0

    def step():
        if args.graphed:
            return graphed_step()
        blocks, x, y = next(it)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.amp):
            logits = model(blocks, x)
        loss = F.cross_entropy(logits.float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss), count_edges(blocks), count_inputs(blocks)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    edges = inputs = 0
    s.record()
    for _ in range(args.steps):
        loss, ne, ni = step()
        edges += ne
        inputs += ni
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    rec = dict(bench=f"{args.model}_uva_pipeline" if args.resident == "host" else f"{args.model}_generic_pipeline_device_resident", impl=args.impl, scale=args.scale, nodes=n, edges=int(u.shape[0]), feat_dim=128,
               batch=args.batch, fanouts=args.fanouts, ms_per_step=ms, sampled_edges_per_s=edges / args.steps / ms * 1e3,
               input_rows_per_step=inputs / args.steps, uva_feature_GBps=inputs / args.steps * 512 / ms / 1e6,
               gpu_cache_mb=args.cache_mb, workers=args.workers, amp=args.amp, graphed=args.graphed, table_input=args.table_input, final_loss=loss, steps=args.steps, warmup=args.warmup, graph_build_s=gen_s)
    print(json.dumps(rec), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(rec) + "\n")
    if hasattr(it, "close"):
        it.close()          # ends the endless batch generator: the loader's producer thread is stopped and joined


if __name__ == "__main__":
    main()
