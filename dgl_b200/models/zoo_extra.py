"""Second half of the model zoo: further model families of the reference's ``examples/pytorch`` tree.

As in ``zoo.py`` every class cites the example it mirrors and is written over this package's own operators: message
passing goes through ``update_all`` / ``dgl_b200.ops`` (g-SpMM, g-SDDMM, edge softmax - the sm_100a kernels on a GPU,
the OpenMP kernels on the host), dense parts are plain ``torch.nn``. Where the reference example carries a custom
layer (ARMA, MixHop, GEN, FiLM, LGNN, CompGCN, GCMC, SAGPool ...) the layer is implemented here, in the formulation
that maps best onto fused aggregation: linear maps are applied on the node side before the sparse step whenever the
algebra allows it (aggregate-after-project), and multi-hop operators reuse one normalised adjacency.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import function as fn
from .. import ops
from ..nn import EdgeConv, GATConv, GMMConv, GraphConv, NNConv, SAGEConv, Set2Set, SortPooling


# ------------------------------------------------------------------------------------------------------ utilities ----
def sym_norm_propagate(g, h, norm=None):
    """One step of D^-1/2 A D^-1/2 h as a single g-SpMM (the scaling is folded into the operands)."""
    if norm is None:
        norm = g.in_degrees().clamp(min=1).to(h.dtype).pow(-0.5).unsqueeze(-1)
    return ops.copy_u_sum(g, h * norm) * norm


def _deg_norm(g, dtype):
    return g.in_degrees().clamp(min=1).to(dtype).pow(-0.5).unsqueeze(-1)


# ------------------------------------------------------------------------------------------------- auto-encoders ----
class GAE(nn.Module):
    """(Variational) graph auto-encoder (examples/pytorch/vgae/model.py): GCN encoder, inner-product decoder.
    ``variational=True`` adds the log-std head and the KL term to :meth:`loss`."""

    def __init__(self, in_size, hid_size, z_size, variational=False):
        super().__init__()
        self.base = GraphConv(in_size, hid_size, activation=F.relu, allow_zero_in_degree=True)
        self.mean = GraphConv(hid_size, z_size, allow_zero_in_degree=True)
        self.log_std = GraphConv(hid_size, z_size, allow_zero_in_degree=True) if variational else None
        self._stats = None

    def encode(self, g, x):
        h = self.base(g, x)
        mu = self.mean(g, h)
        if self.log_std is None:
            return mu
        ls = self.log_std(g, h)
        self._stats = (mu, ls)
        return mu + torch.randn_like(mu) * torch.exp(ls) if self.training else mu

    def decode(self, z, u, v):
        """Logit of edge (u, v)."""
        return (z[u.long()] * z[v.long()]).sum(-1)

    def forward(self, g, x):
        return self.encode(g, x)

    def loss(self, z, pos, neg):
        """Balanced reconstruction loss over positive / negative node pairs (+ KL for the variational model)."""
        lp, ln = self.decode(z, *pos), self.decode(z, *neg)
        rec = F.binary_cross_entropy_with_logits(torch.cat([lp, ln]), torch.cat([torch.ones_like(lp), torch.zeros_like(ln)]))
        if self.log_std is None or self._stats is None:
            return rec
        mu, ls = self._stats
        kl = -0.5 / z.shape[0] * (1 + 2 * ls - mu.pow(2) - torch.exp(2 * ls)).sum(1).mean()
        return rec + kl


# ------------------------------------------------------------------------------------------------- heterogeneous ----
class HAN(nn.Module):
    """Heterogeneous attention network (examples/pytorch/han/model.py): node-level GAT on every meta-path graph, then
    semantic-level attention across meta-paths. ``forward`` takes the list of meta-path reachable graphs."""

    def __init__(self, num_paths, in_size, hid_size, out_size, num_heads=4, dropout=0.0):
        super().__init__()
        self.gats = nn.ModuleList(GATConv(in_size, hid_size, num_heads, dropout, dropout, activation=F.elu,
                                          allow_zero_in_degree=True) for _ in range(num_paths))
        self.semantic = nn.Sequential(nn.Linear(hid_size * num_heads, 128), nn.Tanh(), nn.Linear(128, 1, bias=False))
        self.out = nn.Linear(hid_size * num_heads, out_size)

    @staticmethod
    def metapath_graphs(g, metapaths):
        from ..transforms import metapath_reachable_graph
        return [metapath_reachable_graph(g, mp) for mp in metapaths]

    def forward(self, graphs, x):
        z = torch.stack([gat(g, x).flatten(1) for gat, g in zip(self.gats, graphs)], dim=1)      # [N, P, H*D]
        beta = torch.softmax(self.semantic(z).mean(0), dim=0)                                     # [P, 1]
        return self.out((beta.unsqueeze(0) * z).sum(1))


class GNNFiLM(nn.Module):
    """GNN with feature-wise linear modulation (examples/pytorch/GNN-FiLM/main.py): per relation the target node emits
    (gamma, beta) that modulate the projected source message, messages are summed over relations."""

    class Layer(nn.Module):
        def __init__(self, in_size, out_size, etypes):
            super().__init__()
            self.w = nn.ModuleDict({e: nn.Linear(in_size, out_size, bias=False) for e in etypes})
            self.film = nn.ModuleDict({e: nn.Linear(in_size, 2 * out_size, bias=False) for e in etypes})
            self.norm = nn.LayerNorm(out_size)

        def forward(self, g, feats):
            out = {}
            for s, e, d in g.canonical_etypes:
                if g.num_edges((s, e, d)) == 0:
                    continue
                rel = g[(s, e, d)]
                agg = ops.copy_u_sum(rel, self.w[e](feats[s]))       # sum_u W h_u, then modulate on the node side
                gamma, beta = self.film[e](feats[d]).chunk(2, dim=-1)
                msg = F.relu(gamma * agg + beta * rel.in_degrees().to(agg.dtype).unsqueeze(-1))
                out[d] = out.get(d, 0) + msg
            return {k: self.norm(v) for k, v in out.items()}

    def __init__(self, etypes, in_size, hid_size, out_size, num_layers=2):
        super().__init__()
        sizes = [in_size] + [hid_size] * num_layers
        self.layers = nn.ModuleList(self.Layer(a, b, etypes) for a, b in zip(sizes[:-1], sizes[1:]))
        self.out = nn.Linear(hid_size, out_size)

    def forward(self, g, feats):
        for layer in self.layers:
            new = layer(g, feats)
            feats = {k: new.get(k, torch.zeros(v.shape[0], layer.norm.normalized_shape[0], device=v.device))
                     for k, v in feats.items()}
        return {k: self.out(v) for k, v in feats.items()}


class CompGCN(nn.Module):
    """Composition-based multi-relational GCN (examples/pytorch/compGCN/models.py): messages compose the source
    embedding with the relation embedding (``sub`` / ``mult``); relation embeddings are updated by their own map.
    Runs on a homogeneous graph with an ``etype`` id per edge; edges and their inverses use separate weights."""

    class Layer(nn.Module):
        def __init__(self, in_size, out_size, comp="sub"):
            super().__init__()
            self.comp = comp
            self.w_in, self.w_out, self.w_loop = (nn.Linear(in_size, out_size, bias=False) for _ in range(3))
            self.w_rel = nn.Linear(in_size, out_size, bias=False)
            self.loop_rel = nn.Parameter(torch.randn(1, in_size) * 0.1)

        def compose(self, h, r):
            return h - r if self.comp == "sub" else h * r

        def forward(self, g, h, rel, etype, is_inverse):
            msg = self.compose(h[g.edges()[0].long()], rel[etype.long()])
            msg = torch.where(is_inverse.unsqueeze(-1), self.w_out(msg), self.w_in(msg))
            agg = ops.copy_e_sum(g, msg) / g.in_degrees().clamp(min=1).to(msg.dtype).unsqueeze(-1)
            return torch.tanh((agg + self.w_loop(self.compose(h, self.loop_rel))) / 2), self.w_rel(rel)

    def __init__(self, num_nodes, num_rels, hid_size, out_size, num_layers=2, comp="sub", in_size=None):
        super().__init__()
        # featureless graphs (the knowledge-graph setting of the example) learn one embedding per node; with input
        # features a projection takes its place
        self.node_emb = nn.Parameter(torch.randn(num_nodes, hid_size) * 0.1) if in_size is None else None
        self.inp = None if in_size is None else nn.Linear(in_size, hid_size)
        self.rel_emb = nn.Parameter(torch.randn(2 * num_rels, hid_size) * 0.1)
        self.num_rels = num_rels
        self.layers = nn.ModuleList(self.Layer(hid_size, hid_size, comp) for _ in range(num_layers))
        self.out = nn.Linear(hid_size, out_size)

    def forward(self, g, etype, x=None):
        h, r = (self.node_emb if self.inp is None else self.inp(x)), self.rel_emb
        inverse = etype >= self.num_rels
        for layer in self.layers:
            h, r = layer(g, h, r, etype, inverse)
        return self.out(h)


class GCMC(nn.Module):
    """Graph convolutional matrix completion (examples/pytorch/gcmc/model.py): one graph conv per rating level on the
    user-item bipartite graph (stacked accumulation), dense layer, bilinear decoder over rating classes.
    ``forward(enc_graph, ufeat, ifeat, users, items)`` returns rating logits for the queried pairs."""

    def __init__(self, num_ratings, u_in, i_in, agg_size, out_size, num_basis=2):
        super().__init__()
        self.num_ratings = num_ratings
        per = agg_size // num_ratings
        self.wu = nn.ModuleList(nn.Linear(u_in, per, bias=False) for _ in range(num_ratings))
        self.wi = nn.ModuleList(nn.Linear(i_in, per, bias=False) for _ in range(num_ratings))
        self.fu, self.fi = nn.Linear(per * num_ratings, out_size), nn.Linear(per * num_ratings, out_size)
        self.basis = nn.Parameter(torch.randn(num_basis, out_size, out_size) * 0.1)
        self.combine = nn.Linear(num_basis, num_ratings, bias=False)

    def forward(self, g, ufeat, ifeat, users, items):
        """``g``: heterograph with edge types ``('user', str(r), 'item')`` and ``('item', 'rev-' + str(r), 'user')``."""
        hu, hi = [], []
        for r in range(1, self.num_ratings + 1):
            fw, bw = g[("user", str(r), "item")], g[("item", f"rev-{r}", "user")]
            cu = fw.out_degrees().clamp(min=1).float().pow(-0.5).unsqueeze(-1)
            ci = fw.in_degrees().clamp(min=1).float().pow(-0.5).unsqueeze(-1)
            hi.append(ops.copy_u_sum(fw, self.wu[r - 1](ufeat) * cu) * ci)
            hu.append(ops.copy_u_sum(bw, self.wi[r - 1](ifeat) * ci) * cu)
        zu, zi = self.fu(F.relu(torch.cat(hu, 1))), self.fi(F.relu(torch.cat(hi, 1)))
        a, b = zu[users.long()], zi[items.long()]
        scores = torch.einsum("nd,kde,ne->nk", a, self.basis, b)
        return self.combine(scores)


# -------------------------------------------------------------------------------------- propagation-style models ----
class SIGN(nn.Module):
    """Scalable inception GNN (examples/pytorch/sign/sign.py): K+1 pre-propagated feature matrices, one feed-forward
    branch per hop, concatenation, output MLP. :meth:`precompute` builds the hop features once."""

    def __init__(self, in_size, hid_size, out_size, num_hops=3, dropout=0.0):
        super().__init__()
        self.branches = nn.ModuleList(nn.Sequential(nn.Linear(in_size, hid_size), nn.PReLU(), nn.Dropout(dropout),
                                                    nn.Linear(hid_size, hid_size)) for _ in range(num_hops + 1))
        self.project = nn.Sequential(nn.PReLU(), nn.Dropout(dropout), nn.Linear(hid_size * (num_hops + 1), hid_size),
                                     nn.PReLU(), nn.Linear(hid_size, out_size))

    @staticmethod
    @torch.no_grad()
    def precompute(g, x, num_hops=3):
        norm = _deg_norm(g, x.dtype)
        hops = [x]
        for _ in range(num_hops):
            hops.append(sym_norm_propagate(g, hops[-1], norm))
        return hops

    def forward(self, hops):
        return self.project(torch.cat([b(h) for b, h in zip(self.branches, hops)], dim=-1))


class CorrectAndSmooth(nn.Module):
    """Correct & Smooth post-processing (examples/pytorch/correct_and_smooth/model.py): propagate the residual error
    of the labelled nodes (``autoscale`` or fixed scale), then propagate the corrected predictions with the true
    labels clamped in."""

    def __init__(self, num_correction_layers=20, correction_alpha=0.8, num_smoothing_layers=20, smoothing_alpha=0.8,
                 autoscale=True, scale=1.0):
        super().__init__()
        self.kc, self.ac, self.ks, self.as_ = num_correction_layers, correction_alpha, num_smoothing_layers, smoothing_alpha
        self.autoscale, self.scale = autoscale, scale

    @staticmethod
    def _propagate(g, y, steps, alpha, post=lambda t: t.clamp(0.0, 1.0)):
        norm = _deg_norm(g, y.dtype)
        last = y
        for _ in range(steps):
            y = post(alpha * sym_norm_propagate(g, y, norm) + (1 - alpha) * last)
        return y

    @torch.no_grad()
    def correct(self, g, y_soft, y_true, mask):
        idx = mask.nonzero().squeeze(1) if mask.dtype == torch.bool else mask.long()
        onehot = F.one_hot(y_true.long(), y_soft.shape[1]).to(y_soft.dtype)
        err = torch.zeros_like(y_soft)
        err[idx] = onehot - y_soft[idx]
        if self.autoscale:
            smoothed = self._propagate(g, err, self.kc, self.ac, post=lambda t: t.clamp(-1.0, 1.0))
            sigma = err[idx].abs().sum() / len(idx)
            scale = sigma / smoothed.abs().sum(1, keepdim=True).clamp(min=1e-12)
            scale[scale.isinf() | (scale > 1000)] = 1.0
            out = y_soft + scale * smoothed
            out[out.isnan()] = y_soft[out.isnan()]
            return out

        def fix(t):
            t[idx] = err[idx]
            return t
        return y_soft + self.scale * self._propagate(g, err, self.kc, self.ac, post=fix)

    @torch.no_grad()
    def smooth(self, g, y_soft, y_true, mask):
        idx = mask.nonzero().squeeze(1) if mask.dtype == torch.bool else mask.long()
        y = y_soft.clone()
        y[idx] = F.one_hot(y_true.long(), y_soft.shape[1]).to(y_soft.dtype)
        return self._propagate(g, y, self.ks, self.as_)


class GRAND(nn.Module):
    """Graph random neural network (examples/pytorch/grand/model.py): DropNode, mixed-order propagation
    (mean of A^0..A^K x), MLP; :meth:`consistency_loss` sharpens the average of S augmented predictions."""

    def __init__(self, in_size, hid_size, out_size, order=4, drop_node=0.5, dropout=0.5, samples=2):
        super().__init__()
        self.order, self.drop_node, self.samples = order, drop_node, samples
        self.mlp = nn.Sequential(nn.Dropout(dropout), nn.Linear(in_size, hid_size), nn.ReLU(), nn.Dropout(dropout),
                                 nn.Linear(hid_size, out_size))

    def propagate(self, g, x):
        norm = _deg_norm(g, x.dtype)
        acc, h = x, x
        for _ in range(self.order):
            h = sym_norm_propagate(g, h, norm)
            acc = acc + h
        return acc / (self.order + 1)

    def forward(self, g, x):
        if not self.training:
            return [self.mlp(self.propagate(g, x))]
        outs = []
        for _ in range(self.samples):
            keep = (torch.rand(x.shape[0], 1, device=x.device) >= self.drop_node).to(x.dtype) / (1 - self.drop_node)
            outs.append(self.mlp(self.propagate(g, x * keep)))
        return outs

    @staticmethod
    def consistency_loss(logits, temperature=0.5):
        probs = [torch.softmax(l, -1) for l in logits]
        avg = torch.stack(probs).mean(0)
        sharp = (avg.pow(1.0 / temperature) / avg.pow(1.0 / temperature).sum(-1, keepdim=True)).detach()
        return sum(((p - sharp) ** 2).sum(-1).mean() for p in probs) / len(probs)


class DAGNN(nn.Module):
    """Deep adaptive GNN (examples/pytorch/dagnn/main.py): MLP first, then K propagation steps whose outputs are
    combined by a learned per-node, per-hop retention score."""

    def __init__(self, in_size, hid_size, out_size, k=8, dropout=0.0):
        super().__init__()
        self.k = k
        self.mlp = nn.Sequential(nn.Dropout(dropout), nn.Linear(in_size, hid_size), nn.ReLU(), nn.Dropout(dropout),
                                 nn.Linear(hid_size, out_size))
        self.score = nn.Linear(out_size, 1)

    def forward(self, g, x):
        h = self.mlp(x)
        norm = _deg_norm(g, h.dtype)
        hops = [h]
        for _ in range(self.k):
            hops.append(sym_norm_propagate(g, hops[-1], norm))
        stack = torch.stack(hops, dim=1)                                   # [N, K+1, C]
        gate = torch.sigmoid(self.score(stack))                            # [N, K+1, 1]
        return (gate * stack).sum(1)


class MixHop(nn.Module):
    """MixHop (examples/pytorch/mixhop/mixhop.py): every layer concatenates linear maps of A^p h for p in ``powers``.
    Projection happens before propagation (A^p (h W_p)), so each power costs g-SpMMs at the projected width."""

    class Layer(nn.Module):
        def __init__(self, in_size, out_size, powers):
            super().__init__()
            self.powers = tuple(powers)
            self.lins = nn.ModuleList(nn.Linear(in_size, out_size, bias=False) for _ in self.powers)

        def forward(self, g, h, norm):
            outs = []
            for p, lin in zip(self.powers, self.lins):
                z = lin(h)
                for _ in range(p):
                    z = sym_norm_propagate(g, z, norm)
                outs.append(z)
            return torch.cat(outs, dim=-1)

    def __init__(self, in_size, hid_size, out_size, num_layers=2, powers=(0, 1, 2), dropout=0.0):
        super().__init__()
        sizes = [in_size] + [hid_size * len(powers)] * num_layers
        self.layers = nn.ModuleList(self.Layer(a, hid_size, powers) for a in sizes[:-1])
        self.norms = nn.ModuleList(nn.BatchNorm1d(hid_size * len(powers)) for _ in range(num_layers))
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(hid_size * len(powers), out_size)

    def forward(self, g, x):
        norm = _deg_norm(g, x.dtype)
        h = x
        for layer, bn in zip(self.layers, self.norms):
            h = self.dropout(torch.tanh(bn(layer(g, self.dropout(h), norm))))
        return self.out(h)


class ARMA(nn.Module):
    """ARMA graph filters (examples/pytorch/arma/model.py): K parallel stacks, each T recursive steps
    ``h <- act(L~ h W + x V)``, averaged. Stacks are batched into one weight tensor so a step is one g-SpMM."""

    class Conv(nn.Module):
        def __init__(self, in_size, out_size, num_stacks, num_layers, activation=None, dropout=0.0):
            super().__init__()
            self.k, self.t, self.act = num_stacks, num_layers, activation
            self.w0 = nn.Parameter(torch.empty(num_stacks, in_size, out_size))
            self.w = nn.Parameter(torch.empty(num_stacks, out_size, out_size))
            self.v = nn.Parameter(torch.empty(num_layers, num_stacks, in_size, out_size))
            self.bias = nn.Parameter(torch.zeros(num_layers, num_stacks, 1, out_size))
            self.dropout = nn.Dropout(dropout)
            with torch.no_grad():
                for p in (self.w0, self.w, self.v):
                    for mat in p.view(-1, p.shape[-2], p.shape[-1]):
                        nn.init.xavier_uniform_(mat)

        def forward(self, g, x):
            norm = _deg_norm(g, x.dtype)
            n = x.shape[0]
            h = None
            for t in range(self.t):
                proj = torch.einsum("nf,kfo->nko", x, self.w0) if t == 0 else torch.einsum("nki,kio->nko", h, self.w)
                prop = sym_norm_propagate(g, proj.reshape(n, -1), norm).view(n, self.k, -1)     # all stacks, one SpMM
                skip = torch.einsum("nf,kfo->nko", self.dropout(x), self.v[t])
                h = prop + skip + self.bias[t].transpose(0, 1)
                if self.act is not None:
                    h = self.act(h)
            return h.mean(1)

    def __init__(self, in_size, hid_size, out_size, num_stacks=2, num_layers=2, dropout=0.0):
        super().__init__()
        self.c1 = self.Conv(in_size, hid_size, num_stacks, num_layers, F.relu, dropout)
        self.c2 = self.Conv(hid_size, out_size, num_stacks, num_layers, None, dropout)
        self.dropout = nn.Dropout(dropout)

    def forward(self, g, x):
        return self.c2(g, self.dropout(F.relu(self.c1(g, x))))


class DeeperGCN(nn.Module):
    """DeeperGCN (examples/pytorch/deepergcn/models.py): pre-activation residual blocks of generalised aggregation
    (GENConv): messages ``relu(h_u [+ e]) + eps`` combined by a softmax with learnable temperature ``beta``."""

    class GENConv(nn.Module):
        def __init__(self, size, beta=1.0, learn_beta=True, eps=1e-7):
            super().__init__()
            self.beta = nn.Parameter(torch.tensor(float(beta))) if learn_beta else beta
            self.eps = eps
            self.mlp = nn.Sequential(nn.Linear(size, 2 * size), nn.LayerNorm(2 * size), nn.ReLU(), nn.Linear(2 * size, size))

        def forward(self, g, h, e=None):
            msg = ops.u_add_e(g, h, e) if e is not None else h[g.edges()[0].long()]
            msg = F.relu(msg) + self.eps
            weight = ops.edge_softmax(g, msg * self.beta)
            return self.mlp(h + ops.copy_e_sum(g, msg * weight))

    def __init__(self, in_size, hid_size, out_size, num_layers=4, dropout=0.0):
        super().__init__()
        self.inp = nn.Linear(in_size, hid_size)
        self.convs = nn.ModuleList(self.GENConv(hid_size) for _ in range(num_layers))
        self.norms = nn.ModuleList(nn.LayerNorm(hid_size) for _ in range(num_layers))
        self.dropout = nn.Dropout(dropout)
        self.out = nn.Linear(hid_size, out_size)

    def forward(self, g, x, efeat=None):
        h = self.inp(x)
        h = self.convs[0](g, h, efeat)
        for conv, norm in zip(self.convs[1:], self.norms[1:]):
            h = h + conv(g, self.dropout(F.relu(norm(h))), efeat)
        return self.out(self.dropout(F.relu(self.norms[0](h))))


# --------------------------------------------------------------------------------------------------- contrastive ----
class GRACE(nn.Module):
    """Graph contrastive representation learning (examples/pytorch/grace/model.py): two views by edge dropping and
    feature masking, shared GCN encoder, projection head, symmetric InfoNCE with intra- and inter-view negatives."""

    def __init__(self, in_size, hid_size, proj_size, num_layers=2, tau=0.5):
        super().__init__()
        sizes = [in_size] + [hid_size] * num_layers
        self.convs = nn.ModuleList(GraphConv(a, b, activation=F.relu, allow_zero_in_degree=True)
                                   for a, b in zip(sizes[:-1], sizes[1:]))
        self.proj = nn.Sequential(nn.Linear(hid_size, proj_size), nn.ELU(), nn.Linear(proj_size, hid_size))
        self.tau = tau

    @staticmethod
    def augment(g, x, drop_edge=0.2, mask_feat=0.2):
        from ..transforms import add_self_loop
        u, v = g.edges()
        keep = torch.rand(len(u), device=u.device) >= drop_edge
        from ..convert import graph as make_graph
        view = add_self_loop(make_graph((u[keep], v[keep]), num_nodes=g.num_nodes(), device=g.device))
        cols = (torch.rand(x.shape[1], device=x.device) >= mask_feat).to(x.dtype)
        return view, x * cols

    def forward(self, g, x):
        h = x
        for conv in self.convs:
            h = conv(g, h)
        return h

    def loss(self, z1, z2):
        a, b = F.normalize(self.proj(z1)), F.normalize(self.proj(z2))

        def half(p, q):
            within, across = torch.exp(p @ p.t() / self.tau), torch.exp(p @ q.t() / self.tau)
            return -torch.log(across.diag() / (within.sum(1) + across.sum(1) - within.diag()))
        return 0.5 * (half(a, b) + half(b, a)).mean()


class Node2VecModel(nn.Module):
    """node2vec (examples/pytorch/node2vec/model.py): biased second-order walks as positive contexts, random walks of
    uniform nodes as negatives, skip-gram objective on one embedding table."""

    def __init__(self, g, emb_size=64, walk_length=10, p=1.0, q=1.0, num_walks=2, window_size=5, num_negatives=1):
        super().__init__()
        self.g, self.p, self.q = g, p, q
        self.walk_length, self.num_walks, self.window, self.neg = walk_length, num_walks, window_size, num_negatives
        self.emb = nn.Embedding(g.num_nodes(), emb_size)
        nn.init.normal_(self.emb.weight, std=0.1)

    def sample(self, batch):
        from ..sampling import node2vec_random_walk
        batch = batch.repeat(self.num_walks)
        walks = node2vec_random_walk(self.g, batch, self.p, self.q, self.walk_length - 1)
        windows = walks.unfold(1, self.window, 1).reshape(-1, self.window)
        windows = windows[(windows >= 0).all(1)]
        fake = torch.randint(0, self.g.num_nodes(), (len(batch) * self.neg, self.walk_length))
        fake[:, 0] = batch.repeat(self.neg)
        return windows, fake.unfold(1, self.window, 1).reshape(-1, self.window)

    def forward(self, nodes=None):
        return self.emb.weight if nodes is None else self.emb(nodes)

    def loss(self, pos, neg):
        def score(w):
            z = self.emb(w)
            return (z[:, :1] * z[:, 1:]).sum(-1).reshape(-1)
        return -(F.logsigmoid(score(pos)).mean() + F.logsigmoid(-score(neg)).mean())


# ------------------------------------------------------------------------------------------------ graph-level ----
class MoNet(nn.Module):
    """Gaussian mixture model convolution network (examples/pytorch/monet/citation.py): GMMConv layers over
    degree-based pseudo-coordinates ``(1/sqrt(deg_u), 1/sqrt(deg_v))`` passed through a learned projection."""

    def __init__(self, in_size, hid_size, out_size, num_layers=2, dim=2, n_kernels=3, dropout=0.0):
        super().__init__()
        sizes = [in_size] + [hid_size] * (num_layers - 1) + [out_size]
        self.convs = nn.ModuleList(GMMConv(a, b, dim, n_kernels, allow_zero_in_degree=True) for a, b in zip(sizes[:-1], sizes[1:]))
        self.pseudo = nn.ModuleList(nn.Sequential(nn.Linear(2, dim), nn.Tanh()) for _ in range(num_layers))
        self.dropout = nn.Dropout(dropout)

    @staticmethod
    def pseudo_coordinates(g):
        d = g.in_degrees().clamp(min=1).float().pow(-0.5)
        u, v = g.edges()
        return torch.stack([d[u.long()], d[v.long()]], dim=1)

    def forward(self, g, x, pseudo=None):
        pseudo = self.pseudo_coordinates(g) if pseudo is None else pseudo
        h = x
        for i, (conv, proj) in enumerate(zip(self.convs, self.pseudo)):
            h = conv(g, self.dropout(h), proj(pseudo))
            if i + 1 < len(self.convs):
                h = F.relu(h)
        return h


class EdgeConvNet(nn.Module):
    """Dynamic graph CNN for point clouds (examples/pytorch/pointcloud/edgeconv/model.py): every layer rebuilds the
    k-NN graph in feature space and applies EdgeConv; global max + mean pooling feed the classifier."""

    def __init__(self, in_size, hid_sizes, out_size, k=8):
        super().__init__()
        self.k = k
        sizes = [in_size] + list(hid_sizes)
        self.convs = nn.ModuleList(EdgeConv(a, b, batch_norm=False) for a, b in zip(sizes[:-1], sizes[1:]))
        self.head = nn.Sequential(nn.Linear(2 * sum(hid_sizes), 64), nn.ReLU(), nn.Linear(64, out_size))

    def forward(self, points):
        """``points``: [B, P, C] batch of clouds with the same number of points."""
        from ..transforms import knn_graph
        b, p, _ = points.shape
        h, feats = points, []
        for conv in self.convs:
            g = knn_graph(h, self.k)
            h = conv(g, h.reshape(b * p, -1)).view(b, p, -1)
            feats.append(h)
        cat = torch.cat(feats, dim=-1)
        return self.head(torch.cat([cat.max(1).values, cat.mean(1)], dim=-1))


class MPNN(nn.Module):
    """Neural message passing for molecules (examples/pytorch/... mpnn in the model zoo of the reference docs,
    nn.NNConv + nn.Set2Set): edge-conditioned convolution with a GRU state update and a Set2Set readout."""

    def __init__(self, node_in, edge_in, hid_size=32, out_size=1, steps=3, set2set_iters=3):
        super().__init__()
        self.inp = nn.Linear(node_in, hid_size)
        edge_net = nn.Sequential(nn.Linear(edge_in, 32), nn.ReLU(), nn.Linear(32, hid_size * hid_size))
        self.conv = NNConv(hid_size, hid_size, edge_net, aggregator_type="sum")
        self.gru = nn.GRU(hid_size, hid_size)
        self.steps = steps
        self.readout = Set2Set(hid_size, set2set_iters, 1)
        self.head = nn.Sequential(nn.Linear(2 * hid_size, hid_size), nn.ReLU(), nn.Linear(hid_size, out_size))

    def forward(self, g, x, e):
        h = F.relu(self.inp(x))
        state = h.unsqueeze(0)
        for _ in range(self.steps):
            m = F.relu(self.conv(g, h, e))
            h, state = self.gru(m.unsqueeze(0), state)
            h = h.squeeze(0)
        return self.head(self.readout(g, h))


class DGCNN(nn.Module):
    """Deep graph CNN used by SEAL (examples/pytorch/seal/model.py): GCN stack with tanh, concatenated layer outputs,
    SortPooling to k nodes, 1-D convolutions, dense classifier. :func:`drnl_node_labeling` gives the SEAL node tags."""

    def __init__(self, in_size, hid_size=16, num_layers=3, k=10, out_size=2):
        super().__init__()
        sizes = [in_size] + [hid_size] * num_layers + [1]
        self.convs = nn.ModuleList(GraphConv(a, b, allow_zero_in_degree=True) for a, b in zip(sizes[:-1], sizes[1:]))
        width = hid_size * num_layers + 1
        self.pool = SortPooling(k)
        self.c1 = nn.Conv1d(1, 16, width, width)
        self.c2 = nn.Conv1d(16, 32, min(5, max(1, k // 2)), 1)
        self.k = k
        dense = 32 * (k // 2 - min(5, max(1, k // 2)) + 1)
        self.head = nn.Sequential(nn.Linear(dense, 64), nn.ReLU(), nn.Linear(64, out_size))

    def forward(self, g, x):
        hs, h = [], x
        for conv in self.convs:
            h = torch.tanh(conv(g, h))
            hs.append(h)
        z = self.pool(g, torch.cat(hs, dim=-1)).unsqueeze(1)                 # [B, 1, k * width]
        z = F.relu(self.c1(z))
        z = F.max_pool1d(z, 2, 2)
        z = F.relu(self.c2(z))
        return self.head(z.flatten(1))


def drnl_node_labeling(g, src, dst):
    """Double-radius node labelling of SEAL (examples/pytorch/seal/utils.py:drnl_node_labeling): for the enclosing
    subgraph ``g`` of the target pair, label = 1 + min(ds, dt) + (d // 2) * (d // 2 + d % 2 - 1) with ds / dt the hop
    distances to ``src`` / ``dst`` computed with the other endpoint removed; the endpoints get 1, unreachable nodes 0."""
    n = g.num_nodes()
    u, v = (t.long() for t in g.edges())

    def hops(start, removed):
        keep = (u != removed) & (v != removed)
        uu, vv = u[keep], v[keep]
        dist = torch.full((n,), -1, dtype=torch.long)
        dist[start] = 0
        frontier = torch.zeros(n, dtype=torch.bool)
        frontier[start] = True
        level = 0
        while bool(frontier.any()):
            level += 1
            reach = torch.zeros(n, dtype=torch.bool)
            reach[vv[frontier[uu]]] = True
            reach[uu[frontier[vv]]] = True
            frontier = reach & (dist < 0)
            frontier[removed] = False
            dist[frontier] = level
        return dist

    ds, dt = hops(src, dst), hops(dst, src)
    d = ds + dt
    half, rem = d // 2, d % 2
    z = 1 + torch.minimum(ds, dt) + half * (half + rem - 1)
    z[(ds < 0) | (dt < 0)] = 0
    z[src] = z[dst] = 1
    return z


class SAGPoolNet(nn.Module):
    """Self-attention graph pooling (examples/pytorch/sagpool/network.py, hierarchical variant): GCN -> SAGPool
    (a GCN scores nodes, the top ``ratio`` fraction of every graph is kept and gated by tanh(score)) x L, with
    mean || max readouts summed over levels."""

    class Pool(nn.Module):
        def __init__(self, size, ratio=0.5):
            super().__init__()
            self.score, self.ratio = GraphConv(size, 1, allow_zero_in_degree=True), ratio

        def forward(self, g, h):
            from ..subgraph import node_subgraph
            s = self.score(g, h).squeeze(-1)
            sizes = g.batch_num_nodes()
            gid = torch.repeat_interleave(torch.arange(len(sizes), device=h.device), sizes)
            keep_n = torch.ceil(sizes.float() * self.ratio).long().clamp(min=1)
            order = torch.argsort(gid.double() * 1e6 - s.double().clamp(-1e5, 1e5), stable=True)      # by graph, score desc
            starts = torch.cumsum(sizes, 0) - sizes
            rank = torch.arange(len(order), device=h.device) - starts[gid[order]]
            chosen = torch.sort(order[rank < keep_n[gid[order]]]).values
            sub = node_subgraph(g, chosen)
            sub.set_batch_num_nodes(keep_n)
            ecount = torch.zeros(len(sizes), dtype=torch.long, device=h.device)
            if sub.num_edges():
                ecount.index_add_(0, gid[chosen][sub.edges()[0].long()], torch.ones(sub.num_edges(), dtype=torch.long, device=h.device))
            sub.set_batch_num_edges(ecount)
            return sub, h[chosen] * torch.tanh(s[chosen]).unsqueeze(-1)

    def __init__(self, in_size, hid_size, out_size, num_layers=3, ratio=0.5):
        super().__init__()
        sizes = [in_size] + [hid_size] * num_layers
        self.convs = nn.ModuleList(GraphConv(a, b, activation=F.relu, allow_zero_in_degree=True) for a, b in zip(sizes[:-1], sizes[1:]))
        self.pools = nn.ModuleList(self.Pool(hid_size, ratio) for _ in range(num_layers))
        self.head = nn.Sequential(nn.Linear(2 * hid_size, hid_size), nn.ReLU(), nn.Linear(hid_size, out_size))

    def forward(self, g, x):
        from ..readout import readout_nodes
        h, total = x, 0
        for conv, pool in zip(self.convs, self.pools):
            g, h = pool(g, conv(g, h))
            g.ndata["_h"] = h
            total = total + torch.cat([readout_nodes(g, "_h", op="mean"), readout_nodes(g, "_h", op="max")], dim=-1)
        return self.head(total)


class LGNN(nn.Module):
    """Line graph neural network for community detection (examples/pytorch/line_graph/gnn.py): node features and
    line-graph (edge) features are updated jointly - each side aggregates its own multi-radius neighbourhoods
    (A^(2^j)) and the other side through the incidence structure (``pm`` / ``pd`` in the example)."""

    class Layer(nn.Module):
        def __init__(self, in_size, out_size, radius):
            super().__init__()
            self.radius = radius
            make = lambda: nn.ModuleList(nn.Linear(in_size, out_size) for _ in range(radius + 2))  # noqa: E731
            self.g_self, self.lg_self = make(), make()
            self.g_fuse, self.lg_fuse = nn.Linear(in_size, out_size), nn.Linear(in_size, out_size)
            self.g_bn, self.lg_bn = nn.BatchNorm1d(out_size), nn.BatchNorm1d(out_size)

        @staticmethod
        def _side(graph, h, deg, lins, radius):
            out = lins[0](h) + lins[1](deg * h)
            z, done = h, 0
            for j in range(radius):                      # A^(2^j) h, reusing the hops already taken
                while done < 2 ** j:
                    z = ops.copy_u_sum(graph, z)
                    done += 1
                out = out + lins[2 + j](z)
            return out

        def forward(self, g, lg, x, y, deg_g, deg_lg):
            u, v = (t.long() for t in g.edges())
            # edge features flow to their end points; node features flow to the incident edges
            y_to_x = torch.zeros_like(x[:, : y.shape[1]]).index_add(0, v, y).index_add(0, u, y)
            x_next = self._side(g, x, deg_g, self.g_self, self.radius) + self.g_fuse(y_to_x)
            y_next = self._side(lg, y, deg_lg, self.lg_self, self.radius) + self.lg_fuse(x[u] + x[v])
            half = x_next.shape[1] // 2
            x_next = torch.cat([x_next[:, :half], F.relu(x_next[:, half:])], 1)
            y_next = torch.cat([y_next[:, :half], F.relu(y_next[:, half:])], 1)
            return self.g_bn(x_next), self.lg_bn(y_next)

    def __init__(self, hid_size, out_size, num_layers=3, radius=2):
        super().__init__()
        sizes = [1] + [hid_size] * num_layers
        self.layers = nn.ModuleList(self.Layer(a, b, radius) for a, b in zip(sizes[:-1], sizes[1:]))
        self.out = nn.Linear(hid_size, out_size)

    def forward(self, g, lg):
        deg_g = g.in_degrees().float().unsqueeze(-1)
        deg_lg = lg.in_degrees().float().unsqueeze(-1)
        x, y = deg_g, deg_lg
        for layer in self.layers:
            x, y = layer(g, lg, x, y, deg_g, deg_lg)
        return self.out(x)


# ------------------------------------------------------------------------------------------------ spatio-temporal ----
class STGCN(nn.Module):
    """Spatio-temporal graph convolution (examples/pytorch/stgcn_wave/model.py): blocks of gated temporal convolution
    -> graph convolution -> temporal convolution with layer norm, then an output temporal layer.
    Input ``x``: [B, C, T, N] over the fixed sensor graph ``g``."""

    class Temporal(nn.Module):
        def __init__(self, cin, cout, kernel=3):
            super().__init__()
            self.conv = nn.Conv2d(cin, 2 * cout, (kernel, 1))
            self.res = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()
            self.kernel = kernel

        def forward(self, x):
            a, b = self.conv(x).chunk(2, dim=1)
            return (a + self.res(x)[:, :, self.kernel - 1:, :]) * torch.sigmoid(b)            # GLU

    class Spatial(nn.Module):
        def __init__(self, size):
            super().__init__()
            self.lin = nn.Linear(size, size)

        def forward(self, g, x):
            b, c, t, n = x.shape
            flat = x.permute(3, 0, 2, 1).reshape(n, b * t * c)                                  # nodes first: one g-SpMM
            out = sym_norm_propagate(g, flat).view(n, b, t, c)
            return F.relu(self.lin(out)).permute(1, 3, 2, 0)

    def __init__(self, in_channels, hid_size, num_nodes, window, horizon_out=1, kernel=3):
        super().__init__()
        self.t1, self.s1, self.t2 = self.Temporal(in_channels, hid_size, kernel), self.Spatial(hid_size), self.Temporal(hid_size, hid_size, kernel)
        self.ln1 = nn.LayerNorm([num_nodes, hid_size])
        self.t3, self.s2, self.t4 = self.Temporal(hid_size, hid_size, kernel), self.Spatial(hid_size), self.Temporal(hid_size, hid_size, kernel)
        self.ln2 = nn.LayerNorm([num_nodes, hid_size])
        left = window - 4 * (kernel - 1)
        if left < 1:
            raise ValueError(f"window {window} too short for four temporal convolutions of kernel {kernel}")
        self.out = nn.Sequential(nn.Conv2d(hid_size, hid_size, (left, 1)), nn.ReLU(), nn.Conv2d(hid_size, horizon_out, 1))

    def forward(self, g, x):
        h = self.t2(self.s1(g, self.t1(x)))
        h = self.ln1(h.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        h = self.t4(self.s2(g, self.t3(h)))
        h = self.ln2(h.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return self.out(h).squeeze(2)                                                          # [B, horizon, N]


class PinSAGEModel(nn.Module):
    """PinSAGE item embedding model (examples/pytorch/pinsage/layers.py): importance-weighted SAGE layers over the
    blocks produced by :class:`dgl_b200.sampling.PinSAGESampler` (edge weights = random-walk visit counts), L2-normalised
    outputs, max-margin loss on (query, positive, negative) item triples."""

    class Layer(nn.Module):
        def __init__(self, in_size, hid_size, out_size):
            super().__init__()
            self.q, self.w = nn.Linear(in_size, hid_size), nn.Linear(in_size + hid_size, out_size)

        def forward(self, block, h_src, h_dst, weights):
            w = weights.to(h_src.dtype).unsqueeze(-1)
            agg = ops.u_mul_e_sum(block, F.relu(self.q(h_src)), w)
            total = ops.copy_e_sum(block, w).clamp(min=1)
            z = F.relu(self.w(torch.cat([agg / total, h_dst], dim=1)))
            return z / z.norm(dim=1, keepdim=True).clamp(min=1e-6)

    def __init__(self, in_size, hid_size, num_layers=2):
        super().__init__()
        self.proj = nn.Linear(in_size, hid_size)
        self.layers = nn.ModuleList(self.Layer(hid_size, hid_size, hid_size) for _ in range(num_layers))

    def forward(self, blocks, x, weight_name="weights"):
        h = self.proj(x)
        for layer, block in zip(self.layers, blocks):
            h = layer(block, h, h[: block.num_dst_nodes()], block.edata[weight_name])
        return h

    @staticmethod
    def margin_loss(z_query, z_pos, z_neg, margin=1.0):
        return (((z_query * z_neg).sum(1) - (z_query * z_pos).sum(1)) + margin).clamp(min=0).mean()


__all__ = ["GAE", "HAN", "GNNFiLM", "CompGCN", "GCMC", "SIGN", "CorrectAndSmooth", "GRAND", "DAGNN", "MixHop", "ARMA",
           "DeeperGCN", "GRACE", "Node2VecModel", "MoNet", "EdgeConvNet", "MPNN", "DGCNN", "drnl_node_labeling",
           "SAGPoolNet", "LGNN", "STGCN", "PinSAGEModel", "sym_norm_propagate"]
