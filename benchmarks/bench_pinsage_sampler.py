"""PinSAGE neighbourhood sampler benchmark (random walks with restart + visit counting + top-k), both arms:

  python benchmarks/bench_pinsage_sampler.py --impl ours        # device walk kernel + shared-memory visit counter
  python benchmarks/bench_pinsage_sampler.py --impl ours-sort   # same walks, torch sort / run-length counting (old path)
  python benchmarks/bench_pinsage_sampler.py --impl reference   # unmodified DGL from baseline/_ref

Graph: bipartite item <-> user graph with the shape of the reference's PinSAGE example data after its preprocessing
(examples/pytorch/pinsage: MovieLens-like, scaled up): items x users x interactions, both directions stored.
A step = one call of ``PinSAGESampler(seeds)`` for a batch of item seeds (graph construction of the result included).
CUDA-event time over ``--steps`` calls after ``--warmup``; one JSON line per configuration."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch

CONFIGS = {
    # name: (num_traversals, termination_prob, num_random_walks, num_neighbors)
    "example": (2, 0.5, 10, 3),      # examples/pytorch/pinsage/model.py defaults
    "deep": (5, 0.5, 50, 10),
    "wide": (3, 0.5, 200, 20),
}


def load_impl(impl):
    if impl.startswith("ours"):
        sys.path.insert(0, ROOT)
        import dgl_b200 as dgl
        return dgl
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    os.environ.setdefault("DGLBACKEND", "pytorch")
    import dgl
    return dgl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "ours-sort", "reference"])
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--users", type=int, default=400_000)
    ap.add_argument("--edges", type=int, default=20_000_000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--configs", default="example,deep,wide")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dgl = load_impl(args.impl)
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(0)
    item = torch.randint(0, args.items, (args.edges,), generator=gen)
    user = torch.randint(0, args.users, (args.edges,), generator=gen)
    g = dgl.heterograph({("item", "by", "user"): (item, user), ("user", "buys", "item"): (user, item)},
                        num_nodes_dict={"item": args.items, "user": args.users}).to(dev)
    for name in args.configs.split(","):
        nt, tp, nw, nn = CONFIGS[name]
        sampler = dgl.sampling.PinSAGESampler(g, "item", "user", nt, tp, nw, nn)
        if args.impl == "ours-sort":
            sampler.device_counter = False  # force the sort / run-length counting
        call = sampler
        seeds = [torch.randint(0, args.items, (args.batch,), generator=gen).to(dev) for _ in range(args.steps + args.warmup)]
        for i in range(args.warmup):
            fr = call(seeds[i])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        edges = 0
        for i in range(args.warmup, args.warmup + args.steps):
            fr = call(seeds[i])
            edges += fr.num_edges()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        rec = {"bench": "pinsage_sampler", "impl": args.impl, "config": name, "num_traversals": nt, "termination_prob": tp,
               "num_random_walks": nw, "num_neighbors": nn, "batch": args.batch, "items": args.items, "users": args.users,
               "edges_each_direction": args.edges, "ms_per_call": ms, "seeds_per_s": args.batch / ms * 1e3,
               "neighbors_per_call": edges / args.steps, "steps": args.steps, "warmup": args.warmup}
        print(json.dumps(rec))
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
