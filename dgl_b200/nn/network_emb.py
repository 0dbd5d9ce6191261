"""DeepWalk and MetaPath2Vec embedding models (parity: /root/reference/python/dgl/nn/pytorch/network_emb.py)."""
from __future__ import annotations


import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init

from ..sampling import random_walk


class DeepWalk(nn.Module):
    def __init__(self, g, emb_dim=128, walk_length=40, window_size=5, neg_weight=1, negative_size=5,
                 fast_neg=True, sparse=True):
        super().__init__()
        assert walk_length >= window_size + 1, f"Expect walk_length >= window_size + 1, got {walk_length} and {window_size + 1}"
        self.g = g
        self.emb_dim, self.window_size, self.walk_length = emb_dim, window_size, walk_length
        self.neg_weight, self.negative_size, self.fast_neg = neg_weight, negative_size, fast_neg
        num_nodes = g.num_nodes()
        self.node_embed = nn.Embedding(num_nodes, emb_dim, sparse=sparse)
        self.context_embed = nn.Embedding(num_nodes, emb_dim, sparse=sparse)
        self.reset_parameters()
        if not fast_neg:
            neg_prob = g.out_degrees().pow(0.75)
            self.neg_prob = neg_prob / neg_prob.sum()
        idx_list_src, idx_list_dst = [], []
        for i in range(walk_length):
            for j in range(max(0, i - window_size), i):
                idx_list_src.append(j); idx_list_dst.append(i)
            for j in range(i + 1, min(walk_length, i + 1 + window_size)):
                idx_list_src.append(j); idx_list_dst.append(i)
        self.idx_list_src = torch.LongTensor(idx_list_src)
        self.idx_list_dst = torch.LongTensor(idx_list_dst)

    def reset_parameters(self):
        init_range = 1.0 / self.emb_dim
        init.uniform_(self.node_embed.weight.data, -init_range, init_range)
        init.constant_(self.context_embed.weight.data, 0)

    def sample(self, indices):
        return random_walk(self.g, indices, length=self.walk_length - 1)[0]

    def forward(self, batch_walk):
        """Skip-gram loss of a batch of walks ``[B, walk_length]``: every (centre, context) pair inside the window is a
        positive; each positive gets ``negative_size`` negatives — with ``fast_neg`` they are context vectors of random
        positions of the same batch (no extra embedding lookups), otherwise nodes drawn from degree^0.75. All draws use
        torch's generator, so a seeded run is reproducible."""
        n_walks, length = batch_walk.shape
        dev = batch_walk.device
        centre = self.node_embed(batch_walk)                       # [B, L, D]
        context = self.context_embed(batch_walk)                   # [B, L, D]
        src, dst = self.idx_list_src.to(dev), self.idx_list_dst.to(dev)
        pos = (centre[:, src] * context[:, dst]).sum(-1).clamp(-6, 6)              # [B, P]
        pos_loss = -F.logsigmoid(pos).mean()
        anchor = centre[:, dst].unsqueeze(2)                                        # [B, P, 1, D]
        n_pairs = src.shape[0]
        if self.fast_neg:
            flat = context.reshape(n_walks * length, self.emb_dim)
            pick = torch.randint(0, n_walks * length, (n_walks, n_pairs, self.negative_size), device=dev)
            neg_ctx = flat[pick]                                                    # [B, P, K, D]
        else:
            pick = torch.multinomial(self.neg_prob, n_walks * n_pairs * self.negative_size, replacement=True).to(dev)
            neg_ctx = self.context_embed(pick).view(n_walks, n_pairs, self.negative_size, self.emb_dim)
        neg = (anchor * neg_ctx).sum(-1).clamp(-6, 6)
        neg_loss = -F.logsigmoid(-neg).mean() * self.negative_size * self.neg_weight
        return pos_loss + neg_loss


class MetaPath2Vec(nn.Module):
    def __init__(self, g, metapath, window_size, emb_dim=128, negative_size=5, sparse=True):
        super().__init__()
        assert len(metapath) + 1 >= window_size, f"Expect len(metapath) >= window_size - 1, got {len(metapath)} and {window_size - 1}"
        self.hg = g
        self.emb_dim, self.metapath, self.window_size, self.negative_size = emb_dim, metapath, window_size, negative_size
        self.node_metapath = [g.to_canonical_etype(metapath[0])[0]]
        for etype in metapath:
            self.node_metapath.append(g.to_canonical_etype(etype)[2])
        self.local_to_global_nid = {}
        num_nodes_total = 0
        for ntype in g.ntypes:
            if ntype in self.node_metapath:
                self.local_to_global_nid[ntype] = torch.arange(g.num_nodes(ntype)) + num_nodes_total
                num_nodes_total += g.num_nodes(ntype)
        self.node_embed = nn.Embedding(num_nodes_total, emb_dim, sparse=sparse)
        self.context_embed = nn.Embedding(num_nodes_total, emb_dim, sparse=sparse)
        self.reset_parameters()
        node_frequency = torch.zeros(num_nodes_total)
        for idx in range(g.num_nodes(self.node_metapath[0])):
            traces, _ = random_walk(g, [idx], metapath=metapath)
            for tr in traces.cpu().numpy():
                tr_nids = [self.local_to_global_nid[self.node_metapath[i]][tr[i]] for i in range(len(tr)) if tr[i] >= 0]
                for t in tr_nids:
                    node_frequency[t] += 1
            if idx > 2000:
                break
        neg_prob = node_frequency.pow(0.75)
        self.neg_prob = neg_prob / neg_prob.sum().clamp(min=1e-12)

    def reset_parameters(self):
        init_range = 1.0 / self.emb_dim
        init.uniform_(self.node_embed.weight.data, -init_range, init_range)
        init.constant_(self.context_embed.weight.data, 0)

    def forward(self, pos_u, pos_v, neg_v):
        emb_u = self.node_embed(pos_u)
        emb_v = self.context_embed(pos_v)
        emb_neg_v = self.context_embed(neg_v)
        score = torch.clamp(torch.sum(torch.mul(emb_u, emb_v), dim=1), max=10, min=-10)
        score = -F.logsigmoid(score)
        neg_score = torch.bmm(emb_neg_v, emb_u.unsqueeze(2)).squeeze()
        neg_score = torch.clamp(neg_score, max=10, min=-10)
        neg_score = -torch.sum(F.logsigmoid(-neg_score), dim=1)
        return torch.mean(score + neg_score)

    def sample(self, indices):
        traces, _ = random_walk(self.hg, indices, metapath=self.metapath)
        u_list, v_list = [], []
        for tr in traces.cpu().numpy():
            tr_nids = [int(self.local_to_global_nid[self.node_metapath[i]][tr[i]]) for i in range(len(tr)) if tr[i] >= 0]
            for i, u in enumerate(tr_nids):
                for j, v in enumerate(tr_nids[max(i - self.window_size, 0): i + self.window_size]):
                    if i == j:
                        continue
                    u_list.append(u); v_list.append(v)
        neg_v = torch.multinomial(self.neg_prob, len(u_list) * self.negative_size, replacement=True).view(len(u_list), self.negative_size)
        return torch.LongTensor(u_list), torch.LongTensor(v_list), neg_v
