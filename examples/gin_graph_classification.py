"""GIN graph classification on a synthetic TU-style dataset (reference examples/pytorch/gin/train.py):
batched graphs through GraphDataLoader, sum readout.   python examples/gin_graph_classification.py --device cuda"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.models import GIN


def make_dataset(n=200, seed=0):
    gen = torch.Generator().manual_seed(seed)
    graphs, labels = [], []
    for i in range(n):
        c = i % 2                                      # class 0: sparse ring-like graphs, class 1: denser graphs
        k = int(torch.randint(12, 24, (1,), generator=gen))
        e = k * (2 if c == 0 else 5)
        g = dgl.add_self_loop(dgl.to_bidirected(dgl.graph((torch.randint(0, k, (e,), generator=gen),
                                                           torch.randint(0, k, (e,), generator=gen)), num_nodes=k)))
        g.ndata["x"] = torch.ones(k, 4)
        graphs.append(g)
        labels.append(c)
    return graphs, torch.tensor(labels)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=15)
    a = ap.parse_args()
    graphs, labels = make_dataset()
    from dgl_b200.dataloading import GraphDataLoader
    dl = GraphDataLoader(list(zip(graphs, labels)), batch_size=32, shuffle=True)
    model = GIN(4, 32, 2, num_layers=3).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for ep in range(a.epochs):
        tot = correct = 0
        for bg, y in dl:
            bg, y = bg.to(a.device), y.to(a.device)
            out = model(bg, bg.ndata["x"])
            loss = F.cross_entropy(out, y)
            opt.zero_grad(); loss.backward(); opt.step()
            correct += int((out.argmax(1) == y).sum()); tot += y.numel()
        print(f"epoch {ep:2d} loss {loss.item():.4f} train acc {correct / tot:.3f}")
    return correct / tot


if __name__ == "__main__":
    main()
