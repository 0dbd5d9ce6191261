"""Two "no message passing at training time" recipes on one graph (reference examples/pytorch/sign/ and
correct_and_smooth/): SIGN trains an MLP on pre-propagated features; Correct & Smooth repairs the predictions of a
plain MLP with two label-propagation passes.   python examples/sign_correct_smooth.py [--device cuda]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F
from _planted import accuracy, planted_partition
from dgl_b200.models import SIGN, CorrectAndSmooth


def train(model, fwd, y, mask, epochs, lr=1e-2):
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=5e-4)
    for _ in range(epochs):
        model.train()
        loss = F.cross_entropy(fwd()[mask], y[mask])
        opt.zero_grad(); loss.backward(); opt.step()
    model.eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=80)
    a = ap.parse_args()
    torch.manual_seed(0)
    g, x, y, tr = planted_partition(noise=2.5, device=a.device)
    hops = SIGN.precompute(g, x, num_hops=3)                      # 3 g-SpMMs, once
    sign = SIGN(x.shape[1], 64, 3, num_hops=3, dropout=0.2).to(a.device)
    train(sign, lambda: sign(hops), y, tr, a.epochs)
    with torch.no_grad():
        sign_acc = accuracy(sign(hops), y, ~tr)
    mlp = torch.nn.Sequential(torch.nn.Linear(x.shape[1], 64), torch.nn.ReLU(), torch.nn.Linear(64, 3)).to(a.device)
    train(mlp, lambda: mlp(x), y, tr, a.epochs)
    with torch.no_grad():
        soft = torch.softmax(mlp(x), -1)
    cs = CorrectAndSmooth(20, 0.8, 20, 0.8)
    fixed = cs.smooth(g, cs.correct(g, soft, y[tr], tr), y[tr], tr)
    mlp_acc, cs_acc = accuracy(soft, y, ~tr), accuracy(fixed, y, ~tr)
    print(f"SIGN {sign_acc:.3f} | MLP {mlp_acc:.3f} -> Correct&Smooth {cs_acc:.3f}")
    return sign_acc, mlp_acc, cs_acc


if __name__ == "__main__":
    main()
