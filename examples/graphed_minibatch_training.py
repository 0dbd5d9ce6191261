"""Mini-batch training of a user-defined model with the whole step captured as ONE CUDA graph
(``dgl_b200.dataloading.GraphedTrainer``): neighbour sampling, frontier compaction, feature gather, the model's forward,
the loss, autograd's backward and the optimizer step replay per batch without a single host synchronisation.

The model is ordinary ``dgl.nn`` code; it receives capacity-shaped blocks (ELL neighbour table + device-side row count)
instead of variable-shaped ones. SAGEConv (tcgen05 fused layer), GraphConv, GINConv and GATConv understand them.

  python examples/graphed_minibatch_training.py --model gcn        # needs a CUDA device
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import dgl_b200 as dgl
import dgl_b200.nn as dglnn
from dgl_b200.dataloading import GraphedTrainer


class Net(nn.Module):
    def __init__(self, kind, in_f, hid, classes):
        super().__init__()
        make = {"sage": lambda i, o: dglnn.SAGEConv(i, o, "mean"),
                "gcn": lambda i, o: dglnn.GraphConv(i, o, norm="right", allow_zero_in_degree=True),
                "gin": lambda i, o: dglnn.GINConv(nn.Linear(i, o), "mean"),
                "gat": lambda i, o: dglnn.GATConv(i, o // 4, 4, allow_zero_in_degree=True)}[kind]
        self.layers = nn.ModuleList([make(in_f, hid), make(hid, hid)])
        self.head = nn.Linear(hid, classes)

    def forward(self, blocks, h):
        for layer, block in zip(self.layers, blocks):
            h = layer(block, h)
            h = torch.relu(h.flatten(1) if h.dim() == 3 else h)
        return self.head(h)


def planted_graph(n, classes, feat, edges, homophily=0.85, seed=0):
    g = torch.Generator().manual_seed(seed)
    y = torch.arange(n) % classes
    x = (torch.randn(classes, feat, generator=g) * 1.5)[y] + torch.randn(n, feat, generator=g)
    dst = torch.randint(0, n, (edges,), generator=g)
    same = torch.rand(edges, generator=g) < homophily
    src = torch.where(same, torch.randint(0, n // classes, (edges,), generator=g) * classes + y[dst],
                      torch.randint(0, n, (edges,), generator=g))
    return dgl.graph((src.int(), dst.int()), num_nodes=n, idtype=torch.int32), x, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="gcn", choices=["sage", "gcn", "gin", "gat"])
    ap.add_argument("--nodes", type=int, default=200_000)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("this example needs a CUDA device")
    dev = torch.device("cuda")
    g, x, y = planted_graph(args.nodes, 8, 64, args.nodes * 15)
    g = g.to(dev)
    g.create_formats_()
    model = Net(args.model, 64, 128, 8).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3, capturable=True)      # capturable: no host reads in step()
    trainer = GraphedTrainer(g, x.to(dev), y.to(dev), model, opt, fanouts=[10, 10], batch_size=args.batch)
    train = torch.randperm(args.nodes)[: args.nodes // 2]
    for epoch in range(args.epochs):
        perm = train[torch.randperm(train.numel())]
        torch.cuda.synchronize()
        t0 = time.time()
        for i in range(0, perm.numel() - args.batch + 1, args.batch):
            loss = trainer.step(perm[i:i + args.batch])                       # enqueue only; the loss stays on the device
        torch.cuda.synchronize()
        steps = perm.numel() // args.batch
        print(f"epoch {epoch} loss {float(loss):.4f} {1e3 * (time.time() - t0) / steps:.3f} ms/step "
              f"({trainer.sampled_edges() / 1e6:.1f} M sampled edges so far)")
    return float(loss)


if __name__ == "__main__":
    main()
