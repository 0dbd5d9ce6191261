"""Community detection with the line graph neural network (reference examples/pytorch/line_graph/): supervised on
stochastic-block-model graphs, tested on fresh ones; labels are only defined up to a permutation.
python examples/line_graph_community.py [--device cuda]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F
from _planted import planted_partition
import dgl_b200 as dgl
from dgl_b200.models import LGNN


def sample(seed, device):
    g, _, y, _ = planted_partition(blocks=2, size=40, p_in=0.3, p_out=0.015, seed=seed, device=device)
    g = dgl.remove_self_loop(g)
    return g, dgl.line_graph(g, backtracking=False), y


def perm_invariant(out, y):
    return torch.minimum(F.cross_entropy(out, y), F.cross_entropy(out, 1 - y))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=60)
    a = ap.parse_args()
    torch.manual_seed(0)
    train = [sample(s, a.device) for s in range(8)]
    test = [sample(100 + s, a.device) for s in range(4)]
    model = LGNN(16, 2, num_layers=3, radius=2).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(a.epochs):
        model.train()
        for g, lg, y in train:
            loss = perm_invariant(model(g, lg), y)
            opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    overlaps = []
    with torch.no_grad():
        for g, lg, y in test:
            acc = float((model(g, lg).argmax(1) == y).float().mean())
            overlaps.append(max(acc, 1 - acc))
    score = sum(overlaps) / len(overlaps)
    print(f"mean overlap on unseen graphs {score:.3f}")
    return score


if __name__ == "__main__":
    main()
