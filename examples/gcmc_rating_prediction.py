"""Graph convolutional matrix completion (reference examples/pytorch/gcmc/): predict held-out ratings of a low-rank
user x item matrix from the observed rating graph (one relation per rating level).
python examples/gcmc_rating_prediction.py [--device cuda]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.models import GCMC


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=150)
    a = ap.parse_args()
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    nu, ni, levels = 80, 60, 5
    score = rng.randn(nu, 3) @ rng.randn(ni, 3).T
    cuts = np.quantile(score, np.linspace(0, 1, levels + 1)[1:-1])
    rating = 1 + (score[..., None] > cuts).sum(-1)                                   # 1..5
    seen = rng.rand(nu, ni) < 0.4
    uu, ii = np.nonzero(seen)
    edges = {}
    for r in range(1, levels + 1):
        sel = rating[uu, ii] == r
        edges[("user", str(r), "item")] = (torch.tensor(uu[sel]), torch.tensor(ii[sel]))
        edges[("item", f"rev-{r}", "user")] = (torch.tensor(ii[sel]), torch.tensor(uu[sel]))
    g = dgl.heterograph(edges, {"user": nu, "item": ni}).to(a.device)
    model = GCMC(levels, nu, ni, agg_size=60, out_size=32, num_basis=2).to(a.device)
    uf, itf = torch.eye(nu, device=a.device), torch.eye(ni, device=a.device)
    tu, ti = torch.tensor(uu, device=a.device), torch.tensor(ii, device=a.device)
    target = torch.tensor(rating[uu, ii] - 1, device=a.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(a.epochs):
        loss = F.cross_entropy(model(g, uf, itf, tu, ti), target)
        opt.zero_grad(); loss.backward(); opt.step()
    hu, hi = np.nonzero(~seen)
    with torch.no_grad():
        prob = torch.softmax(model(g, uf, itf, torch.tensor(hu, device=a.device), torch.tensor(hi, device=a.device)), -1)
        expected = (prob * torch.arange(1, levels + 1, device=a.device)).sum(-1).cpu()
    truth = torch.tensor(rating[hu, hi], dtype=torch.float32)
    rmse = float(((expected - truth) ** 2).mean().sqrt())
    base = float(((truth.mean() - truth) ** 2).mean().sqrt())
    print(f"held-out RMSE {rmse:.3f} (predict-the-mean baseline {base:.3f})")
    return rmse, base


if __name__ == "__main__":
    main()
