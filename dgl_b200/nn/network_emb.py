"""Skip-gram node embeddings trained on random walks: DeepWalk (homogeneous) and MetaPath2Vec (typed walks).

Parity targets: /root/reference/python/dgl/nn/pytorch/network_emb.py:17-244 (DeepWalk) and :247-430 (MetaPath2Vec) - same
constructor arguments, parameter names (``node_embed`` / ``context_embed``), ``sample`` collate functions and loss
definitions. Everything here is tensorised: window pairs come from one precomputed (centre, context) offset table, the
MetaPath2Vec vocabulary statistics are one batched walk + ``bincount`` and its ``sample`` emits pairs with a masked gather
instead of per-trace Python loops, so both models can sample and train on the device that holds the graph.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..sampling import random_walk

__all__ = ["DeepWalk", "MetaPath2Vec"]


def _window_offsets(length, window):
    """All ordered position pairs (context, centre) of a walk with 0 < |context - centre| <= window."""
    pos = torch.arange(length)
    gap = pos.view(-1, 1) - pos.view(1, -1)                       # [context, centre]
    ctx, centre = ((gap != 0) & (gap.abs() <= window)).nonzero(as_tuple=True)
    return ctx, centre


class _SkipGramTables(nn.Module):
    """Input ("node") and output ("context") embedding tables with the word2vec initialisation."""

    def __init__(self, num_rows, emb_dim, sparse):
        super().__init__()
        self.emb_dim = emb_dim
        self.node_embed = nn.Embedding(num_rows, emb_dim, sparse=sparse)
        self.context_embed = nn.Embedding(num_rows, emb_dim, sparse=sparse)
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / self.emb_dim
        with torch.no_grad():
            self.node_embed.weight.uniform_(-bound, bound)
            self.context_embed.weight.zero_()


class DeepWalk(_SkipGramTables):
    """DeepWalk: ``sample(seed_nodes)`` walks, ``forward(walks)`` is the negative-sampling skip-gram loss."""

    def __init__(self, g, emb_dim=128, walk_length=40, window_size=5, neg_weight=1, negative_size=5, fast_neg=True,
                 sparse=True):
        if walk_length < window_size + 1:
            raise AssertionError(f"Expect walk_length >= window_size + 1, got {walk_length} and {window_size + 1}")
        super().__init__(g.num_nodes(), emb_dim, sparse)
        self.g = g
        self.walk_length, self.window_size = walk_length, window_size
        self.neg_weight, self.negative_size, self.fast_neg = neg_weight, negative_size, fast_neg
        if not fast_neg:
            unigram = g.out_degrees().to(torch.float32).pow(0.75)
            self.neg_prob = unigram / unigram.sum()
        ctx, centre = _window_offsets(walk_length, window_size)
        self.idx_list_src, self.idx_list_dst = ctx, centre

    def sample(self, indices):
        return random_walk(self.g, indices, length=self.walk_length - 1)[0]

    def forward(self, batch_walk):
        """``batch_walk``: [B, walk_length] node ids. Positives are all in-window pairs; each gets ``negative_size``
        negatives - context rows of random positions of the same batch when ``fast_neg`` (no extra table lookups),
        otherwise nodes drawn from out-degree^0.75. Scores are clamped to [-6, 6] as in word2vec."""
        dev = batch_walk.device
        B, L = batch_walk.shape
        ctx_pos, centre_pos = self.idx_list_src.to(dev), self.idx_list_dst.to(dev)
        inp = self.node_embed(batch_walk)                               # [B, L, D]
        out = self.context_embed(batch_walk)
        pos_score = (inp[:, ctx_pos] * out[:, centre_pos]).sum(-1).clamp(-6, 6)
        P, K = ctx_pos.numel(), self.negative_size
        if self.fast_neg:
            pool = out.reshape(B * L, -1)
            noise = pool[torch.randint(0, B * L, (B, P, K), device=dev)]
        else:
            draws = torch.multinomial(self.neg_prob, B * P * K, replacement=True).to(dev)
            noise = self.context_embed(draws).view(B, P, K, -1)
        neg_score = (inp[:, centre_pos].unsqueeze(2) * noise).sum(-1).clamp(-6, 6)
        return -F.logsigmoid(pos_score).mean() - F.logsigmoid(-neg_score).mean() * (K * self.neg_weight)


class MetaPath2Vec(_SkipGramTables):
    """metapath2vec: walks follow ``metapath``; all node types on the path share one id space
    (``local_to_global_nid[ntype]``). ``sample(start_nodes)`` returns (centre, context, negatives); the loss is the usual
    skip-gram objective. Context = positions within ``window_size`` steps of the centre, excluding itself."""

    _FREQ_WALK_STARTS = 2002           # vocabulary statistics come from walks of (at most) this many start nodes

    def __init__(self, g, metapath, window_size, emb_dim=128, negative_size=5, sparse=True):
        if len(metapath) + 1 < window_size:
            raise AssertionError(f"Expect len(metapath) >= window_size - 1, got {len(metapath)} and {window_size - 1}")
        hops = [g.to_canonical_etype(e) for e in metapath]
        node_metapath = [hops[0][0]] + [h[2] for h in hops]
        on_path = [nt for nt in g.ntypes if nt in node_metapath]
        sizes = [g.num_nodes(nt) for nt in on_path]
        starts = torch.tensor([0] + sizes).cumsum(0)
        super().__init__(int(starts[-1]), emb_dim, sparse)
        self.hg, self.metapath, self.node_metapath = g, metapath, node_metapath
        self.window_size, self.negative_size = window_size, negative_size
        self.local_to_global_nid = {nt: torch.arange(n) + int(s) for nt, n, s in zip(on_path, sizes, starts)}
        # offset of the shared id space at every walk position
        self._pos_offset = torch.tensor([int(self.local_to_global_nid[nt][0]) if g.num_nodes(nt) else 0 for nt in node_metapath])
        self._ctx_pos, self._centre_pos = _window_offsets(len(node_metapath), window_size)
        first = torch.arange(min(g.num_nodes(node_metapath[0]), self._FREQ_WALK_STARTS), device=g.device)
        ids, valid = self._global_traces(first)
        freq = torch.bincount(ids[valid].cpu(), minlength=int(starts[-1])).to(torch.float32).pow(0.75)
        self.neg_prob = freq / freq.sum().clamp(min=1e-12)

    def _global_traces(self, start_nodes):
        """walks as shared-space ids [B, len(node_metapath)] plus the mask of positions reached before a dead end."""
        traces, _ = random_walk(self.hg, start_nodes, metapath=self.metapath)
        valid = traces >= 0
        return traces + self._pos_offset.to(traces.device), valid

    def sample(self, indices):
        ids, valid = self._global_traces(torch.as_tensor(indices))
        ctx, centre = self._ctx_pos.to(ids.device), self._centre_pos.to(ids.device)
        keep = valid[:, ctx] & valid[:, centre]
        pos_u, pos_v = ids[:, centre][keep].cpu(), ids[:, ctx][keep].cpu()
        neg_v = torch.multinomial(self.neg_prob, max(pos_u.numel(), 1) * self.negative_size, replacement=True)
        return pos_u, pos_v, neg_v[:pos_u.numel() * self.negative_size].view(-1, self.negative_size)

    def forward(self, pos_u, pos_v, neg_v):
        centre = self.node_embed(pos_u)                                                    # [P, D]
        pos = (centre * self.context_embed(pos_v)).sum(-1).clamp(-10, 10)
        neg = torch.einsum("pkd,pd->pk", self.context_embed(neg_v), centre).clamp(-10, 10)
        return (-F.logsigmoid(pos) - F.logsigmoid(-neg).sum(-1)).mean()
