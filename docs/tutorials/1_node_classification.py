"""Blitz 1 - node classification with a 2-layer GCN (counterpart of the reference's tutorials/blitz/1_introduction.py).

A graph is built from an edge list, node features / labels live in ``g.ndata``, ``GraphConv`` layers are ordinary
``torch.nn`` modules and the training loop is plain PyTorch."""
import torch
import torch.nn as nn
import torch.nn.functional as F

import dgl_b200 as dgl

dataset = dgl.data.CoraGraphDataset()          # synthetic content with Cora's published shape (no network access)
g = dgl.add_self_loop(dataset[0])
features, labels = g.ndata["feat"], g.ndata["label"]
train_mask, test_mask = g.ndata["train_mask"], g.ndata["test_mask"]


class GCN(nn.Module):
    def __init__(self, in_feats, h_feats, num_classes):
        super().__init__()
        self.conv1 = dgl.nn.GraphConv(in_feats, h_feats)
        self.conv2 = dgl.nn.GraphConv(h_feats, num_classes)

    def forward(self, g, x):
        return self.conv2(g, F.relu(self.conv1(g, x)))


model = GCN(features.shape[1], 16, dataset.num_classes)
opt = torch.optim.Adam(model.parameters(), lr=0.01)
for epoch in range(5):
    logits = model(g, features)
    loss = F.cross_entropy(logits[train_mask], labels[train_mask])
    opt.zero_grad()
    loss.backward()
    opt.step()
    acc = (logits[test_mask].argmax(1) == labels[test_mask]).float().mean()
    print(f"epoch {epoch} loss {loss.item():.3f} test acc {acc:.3f}")
