"""2-layer GCN on a Cora-sized random graph (BASELINE config 1: CPU, world_size=1 plumbing run).
Usage: python examples/gcn_cora.py [--device cuda] [--epochs 50]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.data.synthetic import shaped_dataset
from dgl_b200.models import GCN


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=30)
    args = ap.parse_args()
    g, feat, labels, train_idx, _ = shaped_dataset("cora", device=args.device, train_frac=0.3)
    g = dgl.add_self_loop(g)
    # learnable synthetic task: labels from a random linear map of the aggregated features
    with torch.no_grad():
        w = torch.randn(feat.shape[1], int(labels.max()) + 1, device=feat.device)
        labels = (dgl.ops.copy_u_sum(g, feat) @ w).argmax(1)
    model = GCN(feat.shape[1], 16, int(labels.max()) + 1).to(args.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=5e-4)
    t0 = time.time()
    for epoch in range(args.epochs):
        model.train()
        logits = model(g, feat)
        loss = F.cross_entropy(logits[train_idx.long()], labels[train_idx.long()])
        opt.zero_grad(); loss.backward(); opt.step()
        if epoch % 10 == 0 or epoch == args.epochs - 1:
            model.eval()
            with torch.no_grad():
                acc = (model(g, feat).argmax(1) == labels).float().mean().item()
            print(f"epoch {epoch:3d} loss {loss.item():.4f} acc {acc:.3f}")
    print(f"done in {time.time() - t0:.2f}s")
    return loss.item()


if __name__ == "__main__":
    main()
