"""Spatio-temporal forecasting with STGCN (reference examples/pytorch/stgcn_wave/): sensors on a ring observe a
travelling, decaying wave; the model sees 12 steps and predicts the next one.
python examples/stgcn_forecast.py [--device cuda]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.models import STGCN


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=60)
    a = ap.parse_args()
    torch.manual_seed(0)
    n, window = 20, 12
    ring = dgl.add_self_loop(dgl.to_bidirected(dgl.graph((torch.arange(n), (torch.arange(n) + 1) % n)))).to(a.device)
    t = torch.arange(400).float()
    series = torch.stack([torch.sin(0.15 * t + 2 * torch.pi * i / n) + 0.3 * torch.sin(0.4 * t - 4 * torch.pi * i / n)
                          for i in range(n)], 1) + 0.05 * torch.randn(400, n)
    xs = torch.stack([series[i: i + window] for i in range(380)]).unsqueeze(1).to(a.device)      # [B, 1, T, N]
    ys = torch.stack([series[i + window] for i in range(380)]).to(a.device)
    model = STGCN(1, 16, n, window).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    for _ in range(a.epochs):
        model.train()
        for lo in range(0, 300, 100):
            loss = F.mse_loss(model(ring, xs[lo: lo + 100]).squeeze(1), ys[lo: lo + 100])
            opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    with torch.no_grad():
        mse = float(F.mse_loss(model(ring, xs[300:]).squeeze(1), ys[300:]))
    naive = float(F.mse_loss(xs[300:, 0, -1], ys[300:]))                      # "same as the last step" baseline
    print(f"STGCN test MSE {mse:.4f} (persistence baseline {naive:.4f})")
    return mse, naive


if __name__ == "__main__":
    main()
