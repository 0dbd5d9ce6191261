"""Two label-efficient baselines on one stochastic-block-model graph (reference examples/pytorch/dgi/ and
label_propagation/): Deep Graph Infomax embeddings + logistic-regression probe, and plain label propagation from 10%
of the labels.   python examples/unsupervised_and_propagation.py --device cuda"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.models import DGI, LabelPropagationModel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=60)
    a = ap.parse_args()
    torch.manual_seed(0)
    from dgl_b200.data.sbm import sbm
    blocks, size = 3, 120
    adj = sbm(blocks, size, 10.0, 1.0, np.random.RandomState(2))
    g = dgl.add_self_loop(dgl.graph((torch.as_tensor(adj.row), torch.as_tensor(adj.col)), num_nodes=blocks * size)).to(a.device)
    y = (torch.arange(g.num_nodes()) // size).to(a.device)
    x = (F.one_hot(y, blocks).float() * 0.7 + torch.randn(g.num_nodes(), blocks, device=a.device))
    labelled = torch.rand(g.num_nodes(), device=a.device) < 0.1

    # 1) DGI: no labels during representation learning
    dgi = DGI(blocks, 32).to(a.device)
    opt = torch.optim.Adam(dgi.parameters(), lr=5e-3)
    for _ in range(a.epochs):
        loss = dgi(g, x)
        opt.zero_grad(); loss.backward(); opt.step()
    emb = dgi.embed(g, x)
    clf = torch.nn.Linear(32, blocks).to(a.device)
    copt = torch.optim.Adam(clf.parameters(), lr=5e-2)
    for _ in range(100):
        closs = F.cross_entropy(clf(emb[labelled]), y[labelled])
        copt.zero_grad(); closs.backward(); copt.step()
    dgi_acc = float((clf(emb).argmax(1) == y)[~labelled].float().mean())
    raw = torch.nn.Linear(blocks, blocks).to(a.device)
    ropt = torch.optim.Adam(raw.parameters(), lr=5e-2)
    for _ in range(100):
        rloss = F.cross_entropy(raw(x[labelled]), y[labelled])
        ropt.zero_grad(); rloss.backward(); ropt.step()
    raw_acc = float((raw(x).argmax(1) == y)[~labelled].float().mean())

    # 2) label propagation: no features at all
    lp = LabelPropagationModel(k=20, alpha=0.9)
    lp_acc = float((lp(g, y, labelled).argmax(1) == y)[~labelled].float().mean())
    print(f"raw-feature probe {raw_acc:.3f} | DGI probe {dgi_acc:.3f} | label propagation {lp_acc:.3f}")
    return raw_acc, dgi_acc, lp_acc


if __name__ == "__main__":
    main()
