"""A graph transformer layer written with ``dgl.sparse`` (reference examples/sparse/graph_transformer.py): attention
restricted to the edges of the graph — scores by batched SDDMM, row softmax over the sparse matrix, aggregation by
batched SpMM — plus Laplacian positional encodings. Task: community detection on a stochastic block model.
python examples/sparse_graph_transformer.py --device cuda"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import dgl_b200 as dgl
import dgl_b200.sparse as dglsp


class SparseMHA(nn.Module):
    def __init__(self, hidden, heads):
        super().__init__()
        self.h, self.d = heads, hidden // heads
        self.q, self.k, self.v, self.o = (nn.Linear(hidden, hidden) for _ in range(4))

    def forward(self, A, x):
        n = x.shape[0]
        q = self.q(x).reshape(n, self.d, self.h) * self.d ** -0.5
        k = self.k(x).reshape(n, self.d, self.h)
        v = self.v(x).reshape(n, self.d, self.h)
        attn = dglsp.bsddmm(A, q, k.transpose(1, 0))          # [N, N, heads] on A's sparsity
        attn = attn.softmax()                                  # row-wise over each node's neighbours
        out = dglsp.bspmm(attn, v)                             # [N, d, heads]
        return self.o(out.reshape(n, -1))


class GTLayer(nn.Module):
    def __init__(self, hidden, heads):
        super().__init__()
        self.mha = SparseMHA(hidden, heads)
        self.n1, self.n2 = nn.LayerNorm(hidden), nn.LayerNorm(hidden)
        self.ffn = nn.Sequential(nn.Linear(hidden, 2 * hidden), nn.ReLU(), nn.Linear(2 * hidden, hidden))

    def forward(self, A, h):
        h = self.n1(h + self.mha(A, h))
        return self.n2(h + self.ffn(h))


class GTModel(nn.Module):
    def __init__(self, in_dim, pe_dim, hidden, classes, layers=2, heads=4):
        super().__init__()
        self.inp, self.pe = nn.Linear(in_dim, hidden), nn.Linear(pe_dim, hidden)
        self.layers = nn.ModuleList([GTLayer(hidden, heads) for _ in range(layers)])
        self.out = nn.Linear(hidden, classes)

    def forward(self, A, x, pe):
        h = self.inp(x) + self.pe(pe)
        for layer in self.layers:
            h = layer(A, h)
        return self.out(h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=40)
    a = ap.parse_args()
    torch.manual_seed(0)
    from dgl_b200.data.sbm import sbm
    n_blocks, size = 3, 60
    adj = sbm(n_blocks, size, 14.0, 1.0, np.random.RandomState(0))
    g = dgl.add_self_loop(dgl.graph((torch.as_tensor(adj.row), torch.as_tensor(adj.col)), num_nodes=n_blocks * size))
    y = torch.arange(n_blocks * size) // size
    pe = dgl.lap_pe(g, 8)
    x = torch.randn(g.num_nodes(), 8) * 0.1                    # uninformative features: structure must be used
    u, v = g.edges()
    A = dglsp.from_coo(v, u, shape=(g.num_nodes(), g.num_nodes())).to(a.device)
    model = GTModel(8, 8, 32, n_blocks).to(a.device)
    x, pe, y = x.to(a.device), pe.to(a.device), y.to(a.device)
    train = torch.rand(g.num_nodes(), device=a.device) < 0.3
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    for ep in range(a.epochs):
        out = model(A, x, pe)
        loss = F.cross_entropy(out[train], y[train])
        opt.zero_grad(); loss.backward(); opt.step()
    acc = float((out.argmax(1) == y)[~train].float().mean())
    print(f"loss {loss.item():.4f} held-out acc {acc:.3f}")
    return acc


if __name__ == "__main__":
    main()
