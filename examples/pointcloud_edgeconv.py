"""Dynamic graph CNN on point clouds (reference examples/pytorch/pointcloud/edgeconv/): classify noisy spheres, tori and
cubes; the k-NN graph is rebuilt in feature space before every EdgeConv layer.
python examples/pointcloud_edgeconv.py [--device cuda]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from dgl_b200.models import EdgeConvNet


def cloud(kind, rng, n=48):
    if kind == 0:                                            # sphere surface
        p = rng.randn(n, 3)
        p /= np.linalg.norm(p, axis=1, keepdims=True)
    elif kind == 1:                                          # torus
        a, b = rng.rand(n) * 2 * np.pi, rng.rand(n) * 2 * np.pi
        p = np.stack([(1 + 0.35 * np.cos(b)) * np.cos(a), (1 + 0.35 * np.cos(b)) * np.sin(a), 0.35 * np.sin(b)], 1)
    else:                                                    # cube surface
        p = rng.rand(n, 3) * 2 - 1
        face = rng.randint(0, 3, n)
        p[np.arange(n), face] = np.sign(p[np.arange(n), face])
    rot, _ = np.linalg.qr(rng.randn(3, 3))
    return (p @ rot + 0.02 * rng.randn(n, 3)).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=40)
    a = ap.parse_args()
    torch.manual_seed(0)
    rng = np.random.RandomState(0)
    y = torch.tensor([i % 3 for i in range(150)])
    x = torch.tensor(np.stack([cloud(int(k), rng) for k in y]))
    x, y = x.to(a.device), y.to(a.device)
    model = EdgeConvNet(3, (32, 32), 3, k=6).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    for _ in range(a.epochs):
        model.train()
        for lo in range(0, 120, 40):
            loss = F.cross_entropy(model(x[lo: lo + 40]), y[lo: lo + 40])
            opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    with torch.no_grad():
        acc = float((model(x[120:]).argmax(1) == y[120:]).float().mean())
    print(f"held-out accuracy {acc:.3f} on 30 clouds (3 shapes)")
    return acc


if __name__ == "__main__":
    main()
