"""Distributed GraphSAGE (reference: examples/distributed/graphsage/node_classification.py).
Step 1 (once):   python examples/graphsage_dist.py --partition --num-parts 2 --out /tmp/parts
Step 2 (per job) python tools/launch.py --num_trainers 1 --num_servers 1 --part_config /tmp/parts/g.json \
                      --ip_config ip_config.txt "python examples/graphsage_dist.py --part-config /tmp/parts/g.json"
With no launcher environment it runs in standalone mode (all partitions in this process)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200 import distributed as dd
from dgl_b200.dataloading import NeighborSampler
from dgl_b200.models import GraphSAGE


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--partition", action="store_true")
    ap.add_argument("--num-parts", type=int, default=2)
    ap.add_argument("--out", default="/tmp/dgl_b200_parts")
    ap.add_argument("--part-config", default=None)
    ap.add_argument("--epochs", type=int, default=2)
    args = ap.parse_args()
    if args.partition or args.part_config is None:
        g = dgl.data.rmat_graph(20000, 200000, symmetrize=True, idtype=torch.int64)
        g.ndata["feat"] = torch.randn(g.num_nodes(), 32)
        g.ndata["label"] = torch.randint(0, 5, (g.num_nodes(),))
        g.ndata["train_mask"] = torch.rand(g.num_nodes()) < 0.2
        dd.partition_graph(g, "g", args.num_parts, args.out, num_hops=1, balance_ntypes=g.ndata["train_mask"])
        print("partitioned into", args.out)
        if args.partition:
            return
        args.part_config = os.path.join(args.out, "g.json")
        os.environ.setdefault("DGL_DIST_MODE", "standalone")
    dd.initialize(part_config=args.part_config)
    dg = dd.DistGraph("g")
    train_nids = dd.node_split(dg.ndata["train_mask"], dg.get_partition_book())
    loader = dd.DistNodeDataLoader(dg, train_nids, NeighborSampler([10, 5]), batch_size=512, shuffle=True)
    model = GraphSAGE(32, 64, 5, 2, 0.0)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    for epoch in range(args.epochs):
        tot, nb = 0.0, 0
        for inp, out, blocks in loader:
            x = dg.ndata["feat"][inp]
            y = dg.ndata["label"][out]
            loss = F.cross_entropy(model(blocks, x), y.long())
            opt.zero_grad(); loss.backward(); opt.step()
            tot += loss.item(); nb += 1
        print(f"epoch {epoch} mean loss {tot / max(nb, 1):.4f} over {nb} batches")
    dd.exit_client()


if __name__ == "__main__":
    main()
