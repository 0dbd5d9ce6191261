"""GNNExplainer on BA-Shapes (reference examples/pytorch/gnn_explainer/): train a GCN to recognise the roles of nodes in
attached house motifs, then explain a motif node — edges inside its motif should receive larger mask values than the
Barabasi-Albert background edges of its computation subgraph.   python examples/gnn_explainer.py"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.nn import GNNExplainer, GraphConv


class GCN(nn.Module):
    def __init__(self, i, h, o):
        super().__init__()
        self.convs = nn.ModuleList([GraphConv(i, h, allow_zero_in_degree=True), GraphConv(h, h, allow_zero_in_degree=True),
                                    GraphConv(h, o, allow_zero_in_degree=True)])

    def forward(self, graph, feat, eweight=None):
        h = feat
        for k, conv in enumerate(self.convs):
            h = conv(graph, h, edge_weight=eweight)
            if k < 2:
                h = F.relu(h)
        return h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=300)
    a = ap.parse_args()
    torch.manual_seed(0)
    from dgl_b200.data import BAShapeDataset
    ds = BAShapeDataset(num_base_nodes=120, num_base_edges_per_node=3, num_motifs=24, seed=0)
    g = ds[0]
    y = g.ndata["label"]
    x = torch.cat([g.in_degrees().float().unsqueeze(1) / 10, torch.ones(g.num_nodes(), 1)], 1)
    model = GCN(2, 32, ds.num_classes)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for _ in range(a.epochs):
        loss = F.cross_entropy(model(g, x), y)
        opt.zero_grad(); loss.backward(); opt.step()
    acc = float((model(g, x).argmax(1) == y).float().mean())
    explainer = GNNExplainer(model, num_hops=3, lr=0.05, num_epochs=150)
    ratios = []
    motif_nodes = torch.nonzero(y > 0).view(-1)
    for node in motif_nodes[:: max(1, len(motif_nodes) // 6)][:6].tolist():
        new_center, sg, feat_mask, edge_mask = explainer.explain_node(node, g, x)
        u, v = sg.edges()
        parent = sg.ndata[dgl.NID]
        in_motif = (y[parent[u.long()]] > 0) & (y[parent[v.long()]] > 0)
        if in_motif.any() and (~in_motif).any():
            ratios.append(float(edge_mask[in_motif].mean() / edge_mask[~in_motif].mean().clamp(min=1e-6)))
    ratio = sum(ratios) / max(len(ratios), 1)
    print(f"train accuracy {acc:.3f}; mean(motif-edge mask) / mean(background-edge mask) = {ratio:.2f}")
    return acc, ratio


if __name__ == "__main__":
    main()
