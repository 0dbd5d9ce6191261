"""Model zoo: the model families of the reference's ``examples/pytorch`` tree, built from ``dgl_b200.nn`` layers.

Each class cites the reference example it mirrors. All of them run on CPU and CUDA through the same sparse kernels;
node-level models take ``(g, x)``, graph-level models take a batched graph, link predictors take positive / negative
graphs, hetero models take feature dicts.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import function as fn
from ..nn import (APPNPConv, AvgPooling, ChebConv, EdgePredictor, GATv2Conv, GCN2Conv, GINConv, GraphConv, HeteroGraphConv,
                  HGTConv, JumpingKnowledge, SAGEConv, SGConv, SumPooling, TAGConv)


class SGC(nn.Module):
    """Simplifying GCN (examples/pytorch/sgc/sgc.py): k-hop propagated features + one linear layer."""

    def __init__(self, in_size, out_size, k=2):
        super().__init__()
        self.conv = SGConv(in_size, out_size, k=k, cached=False)

    def forward(self, g, x):
        return self.conv(g, x)


class APPNP(nn.Module):
    """Predict-then-propagate (examples/pytorch/appnp/appnp.py): MLP followed by personalised-PageRank propagation."""

    def __init__(self, in_size, hid_size, out_size, k=10, alpha=0.1, dropout=0.5):
        super().__init__()
        self.lin1, self.lin2 = nn.Linear(in_size, hid_size), nn.Linear(hid_size, out_size)
        self.prop = APPNPConv(k, alpha)
        self.dropout = nn.Dropout(dropout)

    def forward(self, g, x):
        h = self.lin2(self.dropout(F.relu(self.lin1(self.dropout(x)))))
        return self.prop(g, h)


class GIN(nn.Module):
    """Graph isomorphism network for graph classification (examples/pytorch/gin/train.py): GINConv(MLP) stack with
    sum readout of every layer."""

    def __init__(self, in_size, hid_size, out_size, num_layers=4, pooling="sum"):
        super().__init__()
        self.convs, self.bns, self.heads = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for i in range(num_layers):
            mlp = nn.Sequential(nn.Linear(in_size if i == 0 else hid_size, hid_size), nn.ReLU(), nn.Linear(hid_size, hid_size))
            self.convs.append(GINConv(mlp, learn_eps=False))
            self.bns.append(nn.BatchNorm1d(hid_size))
        for i in range(num_layers + 1):
            self.heads.append(nn.Linear(in_size if i == 0 else hid_size, out_size))
        self.pool = SumPooling() if pooling == "sum" else AvgPooling()

    def forward(self, g, x):
        hs = [x]
        h = x
        for conv, bn in zip(self.convs, self.bns):
            h = F.relu(bn(conv(g, h)))
            hs.append(h)
        return sum(head(self.pool(g, h)) for head, h in zip(self.heads, hs))


class GATv2(nn.Module):
    """GATv2 (examples/pytorch/gatv2/train.py)."""

    def __init__(self, in_size, hid_size, out_size, heads=(4, 1)):
        super().__init__()
        self.l1 = GATv2Conv(in_size, hid_size, heads[0], activation=F.elu, allow_zero_in_degree=True)
        self.l2 = GATv2Conv(hid_size * heads[0], out_size, heads[1], allow_zero_in_degree=True)

    def forward(self, g, x):
        return self.l2(g, self.l1(g, x).flatten(1)).mean(1)


class ChebNet(nn.Module):
    """Chebyshev spectral CNN (examples/pytorch/model_zoo/citation_network/models.py ChebNet)."""

    def __init__(self, in_size, hid_size, out_size, k=3):
        super().__init__()
        self.l1, self.l2 = ChebConv(in_size, hid_size, k), ChebConv(hid_size, out_size, k)

    def forward(self, g, x):
        return self.l2(g, F.relu(self.l1(g, x, lambda_max=[2.0])), lambda_max=[2.0])


class TAGCN(nn.Module):
    """Topology-adaptive GCN (examples/pytorch/tagcn/tagcn.py)."""

    def __init__(self, in_size, hid_size, out_size, k=2):
        super().__init__()
        self.l1, self.l2 = TAGConv(in_size, hid_size, k=k, activation=F.relu), TAGConv(hid_size, out_size, k=k)

    def forward(self, g, x):
        return self.l2(g, self.l1(g, x))


class GCNII(nn.Module):
    """Deep GCN with initial residual + identity mapping (examples/pytorch/gcn2/gcn2.py - the reference file name is
    gcnii-style GCN2Conv stack)."""

    def __init__(self, in_size, hid_size, out_size, num_layers=8, alpha=0.1, lambda_=0.5, dropout=0.5):
        super().__init__()
        self.inp, self.out = nn.Linear(in_size, hid_size), nn.Linear(hid_size, out_size)
        self.convs = nn.ModuleList(GCN2Conv(hid_size, layer=i + 1, alpha=alpha, lambda_=lambda_, allow_zero_in_degree=True)
                                   for i in range(num_layers))
        self.dropout = nn.Dropout(dropout)

    def forward(self, g, x):
        h0 = h = F.relu(self.inp(self.dropout(x)))
        for conv in self.convs:
            h = F.relu(conv(g, self.dropout(h), h0))
        return self.out(self.dropout(h))


class JKNet(nn.Module):
    """Jumping-knowledge network (examples/pytorch/jknet/model.py)."""

    def __init__(self, in_size, hid_size, out_size, num_layers=3, mode="cat"):
        super().__init__()
        self.convs = nn.ModuleList(GraphConv(in_size if i == 0 else hid_size, hid_size, activation=F.relu,
                                             allow_zero_in_degree=True) for i in range(num_layers))
        self.jk = JumpingKnowledge(mode, hid_size, num_layers) if mode == "lstm" else JumpingKnowledge(mode)
        self.out = nn.Linear(hid_size * num_layers if mode == "cat" else hid_size, out_size)

    def forward(self, g, x):
        hs, h = [], x
        for conv in self.convs:
            h = conv(g, h)
            hs.append(h)
        return self.out(self.jk(hs))


class LinkPredictor(nn.Module):
    """GraphSAGE encoder + edge scorer trained with negative sampling (examples/pytorch/graphsage/link_pred.py,
    examples/core/link_prediction)."""

    def __init__(self, in_size, hid_size, num_layers=2, scorer="dot"):
        super().__init__()
        self.layers = nn.ModuleList(SAGEConv(in_size if i == 0 else hid_size, hid_size, "mean") for i in range(num_layers))
        self.pred = EdgePredictor(scorer, hid_size, 1) if scorer != "dot" else None

    def encode(self, g, x):
        h = x
        for i, layer in enumerate(self.layers):
            h = layer(g[i] if isinstance(g, (list, tuple)) else g, h)
            if i < len(self.layers) - 1:
                h = F.relu(h)
        return h

    def score(self, pair_graph, h):
        with pair_graph.local_scope():
            pair_graph.ndata["h"] = h
            if self.pred is None:
                pair_graph.apply_edges(fn.u_dot_v("h", "h", "s"))
                return pair_graph.edata["s"].squeeze(-1)
            u, v = pair_graph.edges()
            return self.pred(h[u.long()], h[v.long()]).squeeze(-1)

    def forward(self, g, x, pos_graph, neg_graph):
        h = self.encode(g, x)
        return self.score(pos_graph, h), self.score(neg_graph, h)

    @staticmethod
    def loss(pos, neg):
        return F.binary_cross_entropy_with_logits(torch.cat([pos, neg]), torch.cat([torch.ones_like(pos), torch.zeros_like(neg)]))


class HeteroRGCN(nn.Module):
    """R-GCN on a heterograph with one GraphConv per relation (examples/pytorch/rgcn-hetero/model.py,
    examples/core/rgcn/hetero_rgcn.py)."""

    def __init__(self, in_sizes, hid_size, out_size, rel_names, num_layers=2):
        super().__init__()
        self.proj = nn.ModuleDict({nt: nn.Linear(s, hid_size) for nt, s in in_sizes.items()})
        self.layers = nn.ModuleList(
            HeteroGraphConv({rel: GraphConv(hid_size, out_size if i == num_layers - 1 else hid_size, norm="right",
                                            allow_zero_in_degree=True) for rel in rel_names}, aggregate="sum")
            for i in range(num_layers))

    def forward(self, g, feats):
        h = {nt: F.relu(self.proj[nt](x)) for nt, x in feats.items()}
        for i, layer in enumerate(self.layers):
            h = layer(g, h)
            if i < len(self.layers) - 1:
                h = {k: F.relu(v) for k, v in h.items()}
        return h


class HGT(nn.Module):
    """Heterogeneous graph transformer on the homogenised graph (examples/pytorch/hgt/model.py)."""

    def __init__(self, in_size, hid_size, out_size, num_ntypes, num_etypes, heads=4, num_layers=2):
        super().__init__()
        self.inp = nn.Linear(in_size, hid_size)
        self.layers = nn.ModuleList(HGTConv(hid_size, hid_size // heads, heads, num_ntypes, num_etypes) for _ in range(num_layers))
        self.out = nn.Linear(hid_size, out_size)

    def forward(self, g, x, ntype, etype):
        h = self.inp(x)
        for layer in self.layers:
            h = layer(g, h, ntype, etype)
        return self.out(h)


class DGI(nn.Module):
    """Deep graph infomax (examples/pytorch/dgi/dgi.py): GCN encoder, corruption by feature shuffling, bilinear
    discriminator against the mean summary."""

    def __init__(self, in_size, hid_size):
        super().__init__()
        self.enc = GraphConv(in_size, hid_size, activation=nn.PReLU(hid_size), allow_zero_in_degree=True)
        self.w = nn.Parameter(torch.empty(hid_size, hid_size))
        nn.init.xavier_uniform_(self.w)

    def forward(self, g, x):
        pos = self.enc(g, x)
        neg = self.enc(g, x[torch.randperm(x.shape[0], device=x.device)])
        s = torch.sigmoid(pos.mean(0))
        lp, ln = pos @ (self.w @ s), neg @ (self.w @ s)
        return F.binary_cross_entropy_with_logits(torch.cat([lp, ln]), torch.cat([torch.ones_like(lp), torch.zeros_like(ln)]))

    @torch.no_grad()
    def embed(self, g, x):
        return self.enc(g, x)


class LabelPropagationModel(nn.Module):
    """Correct-and-smooth style label propagation baseline (examples/pytorch/label_propagation)."""

    def __init__(self, k=10, alpha=0.9):
        super().__init__()
        from ..nn import LabelPropagation
        self.lp = LabelPropagation(k, alpha)

    def forward(self, g, labels, mask):
        return self.lp(g, labels, mask)
