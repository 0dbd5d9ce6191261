"""Heterogeneous attention network on an ACM-like paper / author / field graph (reference examples/pytorch/han/): the
meta-paths paper-author-paper and paper-field-paper are materialised once with ``metapath_reachable_graph`` and fused by
semantic attention.   python examples/han_metapath.py [--device cuda]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.models import HAN


def build(seed=0):
    rng = np.random.RandomState(seed)
    n_paper, per_class_authors, n_field = 300, 20, 9
    y = np.repeat(np.arange(3), n_paper // 3)
    pa = [(p, a) for p in range(n_paper)
          for a in (rng.choice(np.arange(y[p] * per_class_authors, (y[p] + 1) * per_class_authors), 2, replace=False)
                    if rng.rand() < 0.8 else rng.randint(0, 3 * per_class_authors, 2))]
    pf = [(p, (y[p] * 3 + rng.randint(0, 3)) if rng.rand() < 0.7 else rng.randint(0, n_field)) for p in range(n_paper)]
    pu, av = map(torch.tensor, zip(*pa))
    fu, fv = map(torch.tensor, zip(*pf))
    hg = dgl.heterograph({("paper", "pa", "author"): (pu, av), ("author", "ap", "paper"): (av, pu),
                          ("paper", "pf", "field"): (fu, fv), ("field", "fp", "paper"): (fv, fu)},
                         {"paper": n_paper, "author": 3 * per_class_authors, "field": n_field})
    return hg, torch.tensor(y)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=80)
    a = ap.parse_args()
    torch.manual_seed(0)
    hg, y = build()
    hg, y = hg.to(a.device), y.to(a.device)
    graphs = HAN.metapath_graphs(hg, [["pa", "ap"], ["pf", "fp"]])
    x = torch.randn(hg.num_nodes("paper"), 16, device=a.device)          # uninformative features: structure must do the work
    train = torch.rand(len(y), device=a.device) < 0.2
    model = HAN(len(graphs), 16, 8, 3, num_heads=4, dropout=0.2).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=1e-3)
    for _ in range(a.epochs):
        model.train()
        loss = F.cross_entropy(model(graphs, x)[train], y[train])
        opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    with torch.no_grad():
        acc = float((model(graphs, x)[~train].argmax(1) == y[~train]).float().mean())
    print(f"HAN test accuracy {acc:.3f} over {len(graphs)} meta-paths")
    return acc


if __name__ == "__main__":
    main()
