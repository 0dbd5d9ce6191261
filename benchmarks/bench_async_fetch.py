"""Overlapped feature fetching benchmark: GraphBolt-style pipeline

    ItemSampler -> sample_neighbor (graph on the GPU) -> fetch_feature (features on DISK behind a host cache) -> train step

with ``overlap_fetch`` off (every read completes before the next stage starts) and on (the FeatureFetcher's software
pipeline over the staged ``read_async`` generators: cache lookup, io_uring row reads and the H2D copy of successive
mini-batches run concurrently on the I/O pool and the copy stream, graphbolt/feature_fetcher.py).

  python benchmarks/bench_async_fetch.py [--nodes 2000000 --dim 128 --cache-mb 256 --steps 60]

One JSON line per mode; CUDA-event time per step including the SAGE forward/backward that consumes the features."""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F

import dgl_b200 as dgl
import dgl_b200.graphbolt as gb
from dgl_b200.graphbolt.impl.features import CPUCachedFeature, DiskBasedFeature


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2_000_000)
    ap.add_argument("--avg-degree", type=int, default=15)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--cache-mb", type=int, default=256)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--fanouts", default="10,10")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--io-threads", type=int, default=8)
    ap.add_argument("--out", default=None)
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()
    dev = torch.device(args.device)
    cuda = dev.type == "cuda"
    gen = torch.Generator().manual_seed(0)
    n, e = args.nodes, args.nodes * args.avg_degree
    g = dgl.graph((torch.randint(0, n, (e,), generator=gen), torch.randint(0, n, (e,), generator=gen)), num_nodes=n)
    graph = gb.from_dglgraph(g).to(dev)
    path = os.path.join(tempfile.mkdtemp(prefix="dglb200_fetch_"), "feat.npy")
    table = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=(n, args.dim))
    rng = np.random.default_rng(0)
    for s in range(0, n, 1 << 18):
        table[s: s + (1 << 18)] = rng.standard_normal((min(1 << 18, n - s), args.dim), dtype=np.float32)
    table.flush()
    del table
    from dgl_b200.graphbolt.internal import futures
    futures.set_num_io_threads(args.io_threads)
    fan = [int(x) for x in args.fanouts.split(",")]
    model = dgl.nn.SAGEConv(args.dim, 256, "mean").to(dev)
    model2 = dgl.nn.SAGEConv(256, 47, "mean").to(dev)
    opt = torch.optim.Adam(list(model.parameters()) + list(model2.parameters()), lr=1e-3)
    labels = torch.randint(0, 47, (n,), generator=gen).to(dev)
    total = (args.steps + args.warmup) * args.batch
    seeds = torch.randint(0, n, (total,), generator=gen)
    for overlap in (False, True):
        disk = DiskBasedFeature(path)
        feat = CPUCachedFeature(disk, args.cache_mb << 20, pin_memory=cuda)
        store = gb.BasicFeatureStore({("node", None, "feat"): feat})
        dp = gb.ItemSampler(gb.ItemSet(seeds, names="seeds"), batch_size=args.batch).copy_to(dev)
        dp = dp.sample_neighbor(graph, fan).fetch_feature(store, node_feature_keys=["feat"], overlap_fetch=overlap)
        import time
        a, b = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if cuda else (None, None)
        step = 0
        h2d = 0
        for mb in dp:
            if step == args.warmup:
                if cuda:
                    torch.cuda.synchronize()
                    a.record()
                t0 = time.perf_counter()
            x = mb.node_features["feat"]
            blocks = mb.blocks
            h = F.relu(model(blocks[0], x))
            out = model2(blocks[1], h)
            loss = F.cross_entropy(out, labels[mb.seeds])
            opt.zero_grad()
            loss.backward()
            opt.step()
            if step >= args.warmup:
                h2d += x.numel() * x.element_size()
            step += 1
        timed = step - args.warmup
        if cuda:
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / max(timed, 1)
        else:
            ms = (time.perf_counter() - t0) * 1e3 / max(timed, 1)
        rec = {"bench": "async_fetch", "overlap_fetch": overlap, "ms_per_step": ms, "steps": timed, "warmup": args.warmup,
               "nodes": n, "edges": e, "dim": args.dim, "cache_mb": args.cache_mb, "batch": args.batch, "fanouts": fan,
               "io_backend": disk.io_backend, "io_threads": args.io_threads, "cache_miss_rate": float(feat.miss_rate),
               "feature_bytes_per_step": h2d / max(timed, 1), "final_loss": float(loss)}
        print(json.dumps(rec))
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
