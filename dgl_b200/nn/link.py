"""Link / edge scoring heads (API parity: /root/reference/python/dgl/nn/pytorch/link/{edgepred,transe,transr}.py)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

_PAIR_OPS = {
    "dot": lambda a, b: (a * b).sum(-1, keepdim=True),
    "cos": lambda a, b: F.cosine_similarity(a, b).unsqueeze(-1),
    "ele": lambda a, b: a * b,
    "cat": lambda a, b: torch.cat((a, b), dim=-1),
}
_PAIR_WIDTH = {"dot": lambda d: 1, "cos": lambda d: 1, "ele": lambda d: d, "cat": lambda d: 2 * d}


class EdgePredictor(nn.Module):
    """Combine the two endpoint embeddings of an edge (dot | cos | ele | cat), optionally followed by a linear map."""

    def __init__(self, op, in_feats=None, out_feats=None, bias=False):
        super().__init__()
        if op not in _PAIR_OPS:
            raise AssertionError(f"Expect op to be in ['dot', 'cos', 'ele', 'cat'], got {op}")
        self.op = op
        project = in_feats is not None and out_feats is not None
        self.linear = nn.Linear(_PAIR_WIDTH[op](in_feats), out_feats, bias=bias) if project else None

    def reset_parameters(self):
        if self.linear is not None:
            self.linear.reset_parameters()

    def forward(self, h_src, h_dst):
        z = _PAIR_OPS[self.op](h_src, h_dst)
        return z if self.linear is None else self.linear(z)


def _translation_score(head, rel, tail, p):
    return -(head + rel - tail).norm(p=p, dim=-1)


class TransE(nn.Module):
    """``-|| h + r - t ||_p`` with one learned translation per relation."""

    def __init__(self, num_rels, feats, p=1):
        super().__init__()
        self.rel_emb = nn.Embedding(num_rels, feats)
        self.p = p

    def reset_parameters(self):
        self.rel_emb.reset_parameters()

    def forward(self, h_head, h_tail, rels):
        return _translation_score(h_head, self.rel_emb(rels), h_tail, self.p)


class TransR(nn.Module):
    """TransE in a relation-specific space: both entities are projected by ``M_r`` [nfeats, rfeats] first."""

    def __init__(self, num_rels, rfeats, nfeats, p=1):
        super().__init__()
        self.rel_emb = nn.Embedding(num_rels, rfeats)
        self.rel_project = nn.Embedding(num_rels, nfeats * rfeats)
        self.rfeats, self.nfeats, self.p = rfeats, nfeats, p

    def reset_parameters(self):
        self.rel_emb.reset_parameters()
        self.rel_project.reset_parameters()

    def forward(self, h_head, h_tail, rels):
        M = self.rel_project(rels).view(-1, self.nfeats, self.rfeats)
        into_rel = lambda h: torch.einsum("nd,ndr->nr", h, M)  # noqa: E731
        return _translation_score(into_rel(h_head), self.rel_emb(rels), into_rel(h_tail), self.p)
