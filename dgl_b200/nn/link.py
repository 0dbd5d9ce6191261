"""Link / edge predictors (parity: /root/reference/python/dgl/nn/pytorch/link/{edgepred,transe,transr}.py)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class EdgePredictor(nn.Module):
    def __init__(self, op, in_feats=None, out_feats=None, bias=False):
        super().__init__()
        assert op in ["dot", "cos", "ele", "cat"], "Expect op to be in ['dot', 'cos', 'ele', 'cat'], got {}".format(op)
        self.op = op
        if (in_feats is not None) and (out_feats is not None):
            if op in ["dot", "cos"]:
                in_feats = 1
            elif op == "cat":
                in_feats = 2 * in_feats
            self.linear = nn.Linear(in_feats, out_feats, bias=bias)
        else:
            self.linear = None

    def reset_parameters(self):
        if self.linear is not None:
            self.linear.reset_parameters()

    def forward(self, h_src, h_dst):
        if self.op == "dot":
            N, D = h_src.shape
            h = torch.bmm(h_src.view(N, 1, D), h_dst.view(N, D, 1)).squeeze(-1)
        elif self.op == "cos":
            h = F.cosine_similarity(h_src, h_dst).unsqueeze(-1)
        elif self.op == "ele":
            h = h_src * h_dst
        else:
            h = torch.cat([h_src, h_dst], dim=-1)
        if self.linear is not None:
            h = self.linear(h)
        return h


class TransE(nn.Module):
    def __init__(self, num_rels, feats, p=1):
        super().__init__()
        self.rel_emb = nn.Embedding(num_rels, feats)
        self.p = p

    def reset_parameters(self):
        self.rel_emb.reset_parameters()

    def forward(self, h_head, h_tail, rels):
        h_rel = self.rel_emb(rels)
        return -torch.norm(h_head + h_rel - h_tail, p=self.p, dim=-1)


class TransR(nn.Module):
    def __init__(self, num_rels, rfeats, nfeats, p=1):
        super().__init__()
        self.rel_emb = nn.Embedding(num_rels, rfeats)
        self.rel_project = nn.Embedding(num_rels, nfeats * rfeats)
        self.rfeats, self.nfeats, self.p = rfeats, nfeats, p

    def reset_parameters(self):
        self.rel_emb.reset_parameters()
        self.rel_project.reset_parameters()

    def forward(self, h_head, h_tail, rels):
        h_rel = self.rel_emb(rels)
        proj_rel = self.rel_project(rels).reshape(-1, self.nfeats, self.rfeats)
        h_head = (h_head.unsqueeze(1) @ proj_rel).squeeze(1)
        h_tail = (h_tail.unsqueeze(1) @ proj_rel).squeeze(1)
        return -torch.norm(h_head + h_rel - h_tail, p=self.p, dim=-1)
