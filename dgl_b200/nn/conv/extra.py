"""TWIRLSUnfoldingAndAttention and GroupRevRes (reference: nn/pytorch/conv/twirlsconv.py:343-555,
nn/pytorch/conv/grouprevres.py)."""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from ... import function as fn


class TWIRLSUnfoldingAndAttention(nn.Module):
    r"""The propagation core of TWIRLS without the MLPs: ``prop_step`` unfolded gradient steps of the graph-regularised
    energy ``||Y - X||^2 + lam * tr(Y^T L Y)`` (optionally pre-conditioned), with IRLS attention re-weighting of the
    edges inserted before step ``attn_bef`` ... of the sequence ``[propagate]*n + [attention] + [propagate]*m``."""

    def __init__(self, d, alp, lam, prop_step, attn_aft=-1, tau=0.2, T=-1, p=1, use_eta=False, init_att=False,
                 attn_dropout=0, precond=True):
        super().__init__()
        self.d, self.lam, self.prop_step, self.precond = d, lam, prop_step, precond
        self.alp = alp if alp > 0 else 1.0 / (lam + 1)
        self.attn_aft, self.tau, self.T, self.p, self.init_att = attn_aft, tau, T, p, init_att
        self.use_eta = use_eta
        self.etas = nn.Parameter(torch.ones(d)) if use_eta else None
        self.attn_dropout = nn.Dropout(attn_dropout)

    def _attention(self, graph, Y):
        graph.ndata["_y"] = Y * self.etas if self.use_eta else Y
        graph.apply_edges(fn.u_sub_v("_y", "_y", "_d"))
        d2 = graph.edata["_d"].square().sum(-1, keepdim=True)
        w = d2.clamp(min=self.tau ** 2) ** (self.p / 2 - 1)
        if self.T > 0:
            w = torch.where(d2 > self.T ** 2, torch.zeros_like(w), w)
        return self.attn_dropout(w)

    def forward(self, g, X):
        with g.local_scope():
            deg = g.in_degrees().to(X).clamp(min=1).unsqueeze(-1)
            w = torch.ones(g.num_edges(), 1, device=X.device, dtype=X.dtype)
            Y = X
            if self.init_att:
                w = self._attention(g, Y)
            for k in range(self.prop_step):
                if k == self.attn_aft and k > 0:
                    w = self._attention(g, Y)
                g.ndata["_y"] = Y
                g.edata["_w"] = w
                g.update_all(fn.u_mul_e("_y", "_w", "m"), fn.sum("m", "_ay"))
                AY = g.ndata["_ay"]
                if self.precond:
                    Y = (1 - self.alp) * Y + self.alp * (self.lam * AY + X) / (self.lam * deg + 1)
                else:
                    Y = (1 - self.alp) * Y + self.alp * (self.lam * AY / deg.sqrt() + X)
            return Y


class GroupRevRes(nn.Module):
    r"""Grouped reversible residual connections (Li et al., "Training Graph Neural Networks with 1000 Layers"):
    the feature dimension is split into ``groups`` chunks ``X_1..X_C``; ``Y_0 = sum_{i>=2} X_i``,
    ``Y_i = f_i(Y_{i-1}) + X_i``. Each group owns a deep copy of ``gnn_module`` (whose in/out size must be
    ``feats // groups``). Activations are recomputed from the outputs in backward when ``torch.utils.checkpoint`` is
    enabled through ``checkpoint=True`` (the reference implements the inverse explicitly, grouprevres.py:11-140)."""

    def __init__(self, gnn_module, groups=2, checkpoint=False):
        super().__init__()
        self.gnn_modules = nn.ModuleList([gnn_module] + [copy.deepcopy(gnn_module) for _ in range(groups - 1)])
        for m in list(self.gnn_modules)[1:]:
            if hasattr(m, "reset_parameters"):
                m.reset_parameters()
        self.groups, self.checkpoint = groups, checkpoint

    def _chunk(self, x):
        return torch.chunk(x, self.groups, dim=-1)

    def forward(self, g, x, *args):
        xs = self._chunk(x)
        extra = [self._chunk(a) if torch.is_tensor(a) and a.shape[-1] == x.shape[-1] and a.shape[0] == x.shape[0] else None
                 for a in args]
        y_in = sum(xs[1:])
        ys = []
        for i in range(self.groups):
            a_i = [e[i] if e is not None else a for e, a in zip(extra, args)]
            if self.checkpoint and y_in.requires_grad:
                from torch.utils.checkpoint import checkpoint
                f = checkpoint(lambda t, i=i, a_i=a_i: self.gnn_modules[i](g, t, *a_i), y_in, use_reentrant=False)
            else:
                f = self.gnn_modules[i](g, y_in, *a_i)
            y_in = f + xs[i]
            ys.append(y_in)
        return torch.cat(ys, dim=-1)
