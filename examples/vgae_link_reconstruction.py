"""Variational graph auto-encoder (reference examples/pytorch/vgae/): hide 15% of the edges, train on the rest, rank the
hidden edges against random non-edges (AUC).   python examples/vgae_link_reconstruction.py [--device cuda] [--plain]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from _planted import planted_partition
import dgl_b200 as dgl
from dgl_b200.models import GAE


def auc(pos, neg):
    return float((pos[:, None] > neg[None, :]).float().mean() + 0.5 * (pos[:, None] == neg[None, :]).float().mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=120)
    ap.add_argument("--plain", action="store_true", help="deterministic GAE instead of the variational model")
    a = ap.parse_args()
    torch.manual_seed(0)
    g, x, _, _ = planted_partition(device=a.device)
    g = dgl.remove_self_loop(g)
    u, v = g.edges()
    once = u < v
    u, v = u[once], v[once]
    perm = torch.randperm(len(u), device=a.device)
    hold, keep = perm[: len(u) * 15 // 100], perm[len(u) * 15 // 100:]
    tg = dgl.add_self_loop(dgl.to_bidirected(dgl.graph((u[keep], v[keep]), num_nodes=g.num_nodes())))
    tu, tv = tg.edges()
    model = GAE(x.shape[1], 64, 32, variational=not a.plain).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    n = g.num_nodes()
    for _ in range(a.epochs):
        model.train()
        z = model(tg, x)
        neg = (torch.randint(0, n, (len(tu),), device=a.device), torch.randint(0, n, (len(tu),), device=a.device))
        loss = model.loss(z, (tu, tv), neg)
        opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    with torch.no_grad():
        z = model(tg, x)
        pos = model.decode(z, u[hold], v[hold])
        neg = model.decode(z, torch.randint(0, n, (2000,), device=a.device), torch.randint(0, n, (2000,), device=a.device))
    score = auc(pos, neg)
    print(f"held-out edge AUC {score:.3f}")
    return score


if __name__ == "__main__":
    main()
