"""SEAL-style link prediction (reference examples/pytorch/seal/): for every candidate pair extract the 1-hop enclosing
subgraph, tag nodes with double-radius labels and classify the subgraph with DGCNN.
python examples/seal_link_prediction.py [--device cuda]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F
from _planted import planted_partition
import dgl_b200 as dgl
from dgl_b200.models import DGCNN, drnl_node_labeling

MAX_LABEL = 12


def enclosing_subgraph(g, a, b):
    """1-hop enclosing subgraph of (a, b) without the target edge, with one-hot DRNL tags as node features."""
    nodes = torch.unique(torch.cat([torch.tensor([a, b]), g.predecessors(a), g.predecessors(b)]))
    sub = dgl.node_subgraph(g, nodes)
    ids = sub.ndata[dgl.NID]
    la, lb = int((ids == a).nonzero()), int((ids == b).nonzero())
    u, v = sub.edges()
    keep = ~(((u == la) & (v == lb)) | ((u == lb) & (v == la)))
    sub = dgl.graph((u[keep], v[keep]), num_nodes=sub.num_nodes())
    z = drnl_node_labeling(sub, la, lb).clamp(max=MAX_LABEL)
    sub.ndata["x"] = F.one_hot(z, MAX_LABEL + 1).float()
    return dgl.add_self_loop(sub)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=25)
    ap.add_argument("--pairs", type=int, default=300)
    a = ap.parse_args()
    torch.manual_seed(0)
    g, _, _, _ = planted_partition(blocks=4, size=40, p_in=0.2, p_out=0.01)
    g = dgl.remove_self_loop(g)
    u, v = g.edges()
    sel = torch.randperm(len(u))[: a.pairs]
    pos = list(zip(u[sel].tolist(), v[sel].tolist()))
    neg = []
    while len(neg) < a.pairs:
        x, y = torch.randint(0, g.num_nodes(), (2,)).tolist()
        if x != y and not g.has_edges_between(x, y):
            neg.append((x, y))
    subs = [enclosing_subgraph(g, x, y) for x, y in pos + neg]
    labels = torch.tensor([1] * len(pos) + [0] * len(neg))
    order = torch.randperm(len(subs))
    cut = int(0.8 * len(subs))
    tr, te = order[:cut], order[cut:]
    model = DGCNN(MAX_LABEL + 1, hid_size=16, num_layers=2, k=10).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)

    def batch(idx):
        bg = dgl.batch([subs[i] for i in idx.tolist()]).to(a.device)
        return bg, bg.ndata["x"], labels[idx].to(a.device)
    for _ in range(a.epochs):
        model.train()
        perm = tr[torch.randperm(len(tr))]
        for lo in range(0, len(perm), 64):
            bg, x, y = batch(perm[lo: lo + 64])
            loss = F.cross_entropy(model(bg, x), y)
            opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    with torch.no_grad():
        bg, x, y = batch(te)
        acc = float((model(bg, x).argmax(1) == y).float().mean())
    print(f"SEAL / DGCNN link classification accuracy {acc:.3f} on {len(te)} held-out pairs")
    return acc


if __name__ == "__main__":
    main()
