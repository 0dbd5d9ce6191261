"""GRACE: contrastive node representations from two corrupted views (reference examples/pytorch/grace/), evaluated with
a linear probe.   python examples/grace_contrastive.py [--device cuda]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F
from _planted import accuracy, planted_partition
from dgl_b200.models import GRACE


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=80)
    a = ap.parse_args()
    torch.manual_seed(0)
    g, x, y, tr = planted_partition(device=a.device)
    model = GRACE(x.shape[1], 64, 64, tau=0.5).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=1e-5)
    for _ in range(a.epochs):
        (g1, x1), (g2, x2) = GRACE.augment(g, x, 0.3, 0.3), GRACE.augment(g, x, 0.4, 0.2)
        loss = model.loss(model(g1, x1), model(g2, x2))
        opt.zero_grad(); loss.backward(); opt.step()
    with torch.no_grad():
        z = F.normalize(model(g, x))
    probe = torch.nn.Linear(z.shape[1], 3).to(a.device)
    popt = torch.optim.Adam(probe.parameters(), lr=5e-2)
    for _ in range(150):
        ploss = F.cross_entropy(probe(z[tr]), y[tr])
        popt.zero_grad(); ploss.backward(); popt.step()
    raw = torch.nn.Linear(x.shape[1], 3).to(a.device)
    ropt = torch.optim.Adam(raw.parameters(), lr=5e-2)
    for _ in range(150):
        rloss = F.cross_entropy(raw(x[tr]), y[tr])
        ropt.zero_grad(); rloss.backward(); ropt.step()
    with torch.no_grad():
        emb_acc, raw_acc = accuracy(probe(z), y, ~tr), accuracy(raw(x), y, ~tr)
    print(f"linear probe: GRACE embeddings {emb_acc:.3f} vs raw features {raw_acc:.3f}")
    return emb_acc, raw_acc


if __name__ == "__main__":
    main()
