"""8-head GAT, full-graph training (BASELINE config 3: Reddit-shaped random graph, g-SDDMM u_dot_v / edge-softmax path).
Usage: python examples/gat_fullgraph.py --device cuda --scale 1.0"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.data.synthetic import shaped_dataset
from dgl_b200.models import GAT


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--dataset", default="cora")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--hidden", type=int, default=8)
    ap.add_argument("--heads", type=int, default=8)
    args = ap.parse_args()
    g, feat, labels, train_idx, _ = shaped_dataset(args.dataset, device=args.device, scale=args.scale, train_frac=0.3)
    g = dgl.add_self_loop(g)
    C = int(labels.max()) + 1
    model = GAT(feat.shape[1], args.hidden, C, heads=(args.heads, 1)).to(args.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3, weight_decay=5e-4)
    for epoch in range(args.epochs):
        if args.device != "cpu":
            torch.cuda.synchronize()
        t0 = time.time()
        logits = model(g, feat)
        loss = F.cross_entropy(logits[train_idx.long()], labels[train_idx.long()].long())
        opt.zero_grad(); loss.backward(); opt.step()
        if args.device != "cpu":
            torch.cuda.synchronize()
        print(f"epoch {epoch:3d} loss {loss.item():.4f} time {1e3 * (time.time() - t0):.1f} ms")
    return loss.item()


if __name__ == "__main__":
    main()
