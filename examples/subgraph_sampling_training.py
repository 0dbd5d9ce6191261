"""Subgraph-sampling training with ClusterGCN and GraphSAINT samplers (reference examples/pytorch/cluster_gcn/,
graphsaint/): each step trains a GCN on an induced subgraph instead of a neighbourhood expansion.
python examples/subgraph_sampling_training.py --sampler saint --device cuda"""
import argparse, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.nn import GraphConv


class GCN(nn.Module):
    def __init__(self, i, h, o):
        super().__init__()
        self.c1, self.c2 = GraphConv(i, h, allow_zero_in_degree=True), GraphConv(h, o, allow_zero_in_degree=True)

    def forward(self, g, x):
        return self.c2(g, F.relu(self.c1(g, x)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--sampler", choices=["cluster", "saint"], default="cluster")
    ap.add_argument("--epochs", type=int, default=30)
    a = ap.parse_args()
    torch.manual_seed(0)
    from dgl_b200.data.sbm import sbm
    blocks, size = 4, 150
    adj = sbm(blocks, size, 12.0, 1.5, np.random.RandomState(1))
    g = dgl.graph((torch.as_tensor(adj.row), torch.as_tensor(adj.col)), num_nodes=blocks * size)
    y = torch.arange(g.num_nodes()) // size
    g.ndata["label"] = y
    g.ndata["feat"] = F.one_hot(y, blocks).float() * 0.8 + torch.randn(g.num_nodes(), blocks)   # weak per-node signal: a
    # single node is right ~45% of the time, aggregation over ~12 same-community neighbours makes it reliable
    g.ndata["train"] = torch.rand(g.num_nodes()) < 0.5
    with tempfile.TemporaryDirectory() as d:
        if a.sampler == "cluster":
            sampler = dgl.dataloading.ClusterGCNSampler(g, 12, cache_path=os.path.join(d, "cluster.pkl"))
            loader = dgl.dataloading.DataLoader(g, torch.arange(12), sampler, batch_size=3, shuffle=True)
        else:
            sampler = dgl.dataloading.SAINTSampler("node", 300)
            loader = dgl.dataloading.DataLoader(g, torch.arange(8), sampler, batch_size=1)
        model = GCN(blocks, 32, blocks).to(a.device)
        opt = torch.optim.Adam(model.parameters(), lr=1e-2)
        for ep in range(a.epochs):
            for sg in loader:
                sg = dgl.add_self_loop(sg.to(a.device))
                out = model(sg, sg.ndata["feat"])
                m = sg.ndata["train"]
                if int(m.sum()) == 0:
                    continue
                loss = F.cross_entropy(out[m], sg.ndata["label"][m])
                opt.zero_grad(); loss.backward(); opt.step()
        model.eval()
        fg = dgl.add_self_loop(g).to(a.device)
        with torch.no_grad():
            pred = model(fg, fg.ndata["feat"]).argmax(1).cpu()
        acc = float((pred == y)[~g.ndata["train"]].float().mean())
    print(f"{a.sampler}: held-out accuracy {acc:.3f}")
    return acc


if __name__ == "__main__":
    main()
