"""Post-hoc explanation of trained GNNs: mask optimisation (GNNExplainer), a parameterised mask predictor (PGExplainer)
and Shapley-scored subgraph search (SubgraphX), each for homogeneous and heterogeneous graphs.

Behavioural parity targets: /root/reference/python/dgl/nn/pytorch/explain/gnnexplainer.py:16-440 (homogeneous),
:442-900 (hetero), pgexplainer.py:13-607 / :609-1140, subgraphx.py:18-368 / :370-760. The organisation is this repo's own:

* the homogeneous explainers are the one-type case of the typed ones - masks, penalties and the optimisation loop live in
  ``_TypedMasks`` / ``_optimise`` and are shared;
* reverse-edge lookups (mask symmetrisation) are a sort + searchsorted on packed (u, v) keys instead of per-edge
  ``edge_ids`` queries;
* SubgraphX draws all coalitions of a Shapley estimate at once as a boolean matrix and can evaluate them as ONE batched
  forward over replicated copies of the graph (``coalition_batch``), the natural shape for a large GPU; connected
  components are a label-propagation fix point on the device.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..base import EID, ETYPE, NID, NTYPE
from ..subgraph import khop_in_subgraph, node_subgraph

__all__ = ["GNNExplainer", "HeteroGNNExplainer", "PGExplainer", "HeteroPGExplainer", "SubgraphX", "HeteroSubgraphX",
           "MCTSNode"]

_EPS = 1e-15


def _bernoulli_entropy(p):
    return -(p * torch.log(p + _EPS) + (1.0 - p) * torch.log(1.0 - p + _EPS))


class _TypedMasks:
    """Learnable logits of one feature mask per node type ((1, D_t)) and one edge mask per relation ((E_r,))."""

    def __init__(self, graph, feats):
        dev = graph.device
        relu_gain = nn.init.calculate_gain("relu")
        self.feat = {nt: nn.Parameter(0.1 * torch.randn(1, x.shape[1], device=dev)) for nt, x in feats.items()}
        self.edge = {}
        for rel in graph.canonical_etypes:
            endpoints = graph.num_nodes(rel[0]) + graph.num_nodes(rel[2])
            spread = relu_gain * math.sqrt(2.0 / max(endpoints, 1))
            self.edge[rel] = nn.Parameter(spread * torch.randn(graph.num_edges(rel), device=dev))

    def parameters(self):
        return [*self.feat.values(), *self.edge.values()]

    def feat_gates(self):
        return {k: m.sigmoid() for k, m in self.feat.items()}

    def edge_gates(self):
        return {k: m.sigmoid() for k, m in self.edge.items()}

    @staticmethod
    def penalty(feat_logits, edge_logits, a_size, a_ent, b_size, b_ent):
        """size + entropy regularisers (sum of gates for edges, mean for features; mean Bernoulli entropy for both)."""
        total = 0.0
        for m in edge_logits:
            g = m.sigmoid()
            total = total + a_size * g.sum() + a_ent * _bernoulli_entropy(g).mean()
        for m in feat_logits:
            g = m.sigmoid()
            total = total + b_size * g.mean() + b_ent * _bernoulli_entropy(g).mean()
        return total


class _MaskExplainerBase(nn.Module):
    def __init__(self, model, num_hops, lr=0.01, num_epochs=100, *, alpha1=0.005, alpha2=1.0, beta1=1.0, beta2=0.1, log=True):
        super().__init__()
        self.model, self.num_hops, self.lr, self.num_epochs, self.log = model, num_hops, lr, num_epochs, log
        self.alpha1, self.alpha2, self.beta1, self.beta2 = alpha1, alpha2, beta1, beta2

    def _coeffs(self):
        return self.alpha1, self.alpha2, self.beta1, self.beta2

    def _prepare(self, graph):
        self.model = self.model.to(graph.device)
        self.model.eval()

    def _optimise(self, masks, forward, objective):
        """Adam on the mask logits: ``forward(feat_gates, edge_gates) -> logits``; ``objective(log_probs) -> scalar``."""
        opt = torch.optim.Adam(masks.parameters(), lr=self.lr)
        for _ in range(self.num_epochs):
            opt.zero_grad()
            logp = forward(masks.feat_gates(), masks.edge_gates()).log_softmax(dim=-1)
            loss = objective(logp) + masks.penalty(masks.feat.values(), masks.edge.values(), *self._coeffs())
            loss.backward()
            opt.step()
        with torch.no_grad():
            return {k: g.reshape(-1) for k, g in masks.feat_gates().items()}, masks.edge_gates()


class HeteroGNNExplainer(_MaskExplainerBase):
    """Soft feature / edge masks per type that keep the model's prediction (``model(graph, feat, eweight=...)`` with dict
    inputs; node-level models return a dict of logits per node type)."""

    def _init_masks(self, graph, feat):
        m = _TypedMasks(graph, feat)
        return m.feat, m.edge

    def _loss_regularize(self, loss, feat_masks, edge_masks):
        return loss + _TypedMasks.penalty(feat_masks.values(), edge_masks.values(), *self._coeffs())

    def explain_graph(self, graph, feat, **kwargs):
        self._prepare(graph)
        with torch.no_grad():
            target = self.model(graph=graph, feat=feat, **kwargs).argmax(dim=-1)
        run = lambda fg, eg: self.model(graph=graph, feat={nt: feat[nt] * fg[nt] for nt in feat}, eweight=eg, **kwargs)
        return self._optimise(_TypedMasks(graph, feat), run, lambda logp: -logp[0, target[0]])

    def explain_node(self, ntype, node_id, graph, feat, **kwargs):
        self._prepare(graph)
        sg, centre = khop_in_subgraph(graph, {ntype: node_id}, self.num_hops)
        centre = centre[ntype]
        local = {nt: feat[nt][sg.nodes[nt].data[NID].long()] for nt in feat}
        with torch.no_grad():
            target = self.model(graph=sg, feat=local, **kwargs)[ntype].argmax(dim=-1)
        run = lambda fg, eg: self.model(graph=sg, feat={nt: local[nt] * fg[nt] for nt in local}, eweight=eg, **kwargs)[ntype]
        fm, em = self._optimise(_TypedMasks(sg, local), run, lambda logp: -logp[centre, target[centre]].sum())
        return centre, sg, fm, em


class GNNExplainer(_MaskExplainerBase):
    """The single-type case: tensors in, tensors out (``model(graph, feat, eweight=...)``)."""

    @staticmethod
    def _typed(graph, feat):
        return {graph.ntypes[0]: feat}

    def _init_masks(self, graph, feat):
        m = _TypedMasks(graph, self._typed(graph, feat))
        return m.feat[graph.ntypes[0]], m.edge[graph.canonical_etypes[0]]

    def _loss_regularize(self, loss, feat_mask, edge_mask):
        return loss + _TypedMasks.penalty([feat_mask], [edge_mask], *self._coeffs())

    def _run(self, graph, feat, objective, **kwargs):
        nt, rel = graph.ntypes[0], graph.canonical_etypes[0]
        fwd = lambda fg, eg: self.model(graph=graph, feat=feat * fg[nt], eweight=eg[rel], **kwargs)
        fm, em = self._optimise(_TypedMasks(graph, self._typed(graph, feat)), fwd, objective)
        return fm[nt], em[rel]

    def explain_graph(self, graph, feat, **kwargs):
        self._prepare(graph)
        with torch.no_grad():
            target = self.model(graph=graph, feat=feat, **kwargs).argmax(dim=-1)
        return self._run(graph, feat, lambda logp: -logp[0, target[0]], **kwargs)

    def explain_node(self, node_id, graph, feat, **kwargs):
        self._prepare(graph)
        sg, centre = khop_in_subgraph(graph, node_id, self.num_hops)
        rows = sg.ndata[NID].long()
        n_full = graph.num_nodes()
        kwargs = {k: (v[rows] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n_full else v) for k, v in kwargs.items()}
        local = feat[rows]
        with torch.no_grad():
            target = self.model(graph=sg, feat=local, **kwargs).argmax(dim=-1)
        fm, em = self._run(sg, local, lambda logp: -logp[centre, target[centre]].sum(), **kwargs)
        return centre, sg, fm, em


# ------------------------------------------------------------------------------------------------------ PGExplainer
def _reverse_values(src, dst, num_nodes, values):
    """values of the reverse edge (v, u) for every edge (u, v); an edge without a reverse keeps its own value."""
    src, dst = src.long(), dst.long()
    fwd, bwd = src * num_nodes + dst, dst * num_nodes + src
    order = torch.argsort(fwd)
    sorted_keys = fwd[order]
    hit = torch.searchsorted(sorted_keys, bwd).clamp_(max=max(sorted_keys.numel() - 1, 0))
    found = sorted_keys[hit] == bwd
    return torch.where(found, values[order[hit]], values)


class PGExplainer(nn.Module):
    """An MLP over (source embedding, destination embedding[, target-node embedding]) scores every edge; scores are relaxed
    to (0, 1) gates with the concrete distribution during training. The explained model is called as
    ``model(graph, feat, embed=bool, edge_weight=...)`` and returns node embeddings when ``embed`` is set."""

    def __init__(self, model, num_features, num_hops=None, explain_graph=True, coff_budget=0.01, coff_connect=5e-4,
                 sample_bias=0.0):
        super().__init__()
        self.model = model
        self.graph_explanation = bool(explain_graph)
        self.num_hops = num_hops
        self.num_features = num_features * (2 if self.graph_explanation else 3)
        self.coff_budget, self.coff_connect, self.sample_bias = coff_budget, coff_connect, sample_bias
        self.init_bias = 0.0
        self.elayers = nn.Sequential(nn.Linear(self.num_features, 64), nn.ReLU(), nn.Linear(64, 1))
        self.edge_mask = None

    # -- bookkeeping kept for API compatibility
    def set_masks(self, graph, edge_mask=None):
        self.edge_mask = edge_mask

    def clear_masks(self):
        self.edge_mask = None

    def parameters(self):  # only the explanation network is trained
        return self.elayers.parameters()

    def _check_mode(self, graph_level):
        want = "True" if graph_level else "False"
        assert self.graph_explanation == graph_level, f'"explain_graph" must be {want} when initializing the module.'
        if not graph_level:
            assert self.num_hops is not None, '"num_hops" must be provided when initializing the module.'

    def _to(self, device):
        self.model = self.model.to(device)
        self.elayers = self.elayers.to(device)

    # -- objective
    def loss(self, prob, ori_pred):
        nll = -torch.log(prob.gather(-1, ori_pred.unsqueeze(-1)) + 1e-6).mean()
        gates = self.sparse_mask_values
        budget = gates.sum() if self.coff_budget <= 0 else F.relu(gates.sum() - self.coff_budget)
        squeezed = self.edge_mask * 0.98 + 0.01      # keep the entropy finite at saturated gates
        connect = -(squeezed * squeezed.log() + (1 - squeezed) * (1 - squeezed).log()).mean()
        return nll + self.coff_budget * budget + self.coff_connect * connect

    def concrete_sample(self, w, beta=1.0, training=True):
        if not training:
            return w.sigmoid()
        u = self.sample_bias + (1.0 - 2.0 * self.sample_bias) * torch.rand_like(w)
        return ((u.log() - (1.0 - u).log() + w) / beta).sigmoid()

    def _gate(self, pair_embeddings, temperature, training):
        gates = self.concrete_sample(self.elayers(pair_embeddings).reshape(-1), beta=temperature, training=training)
        self.sparse_mask_values = gates
        return gates

    def _finish(self, probs, training):
        if not training:
            self.clear_masks()
        return probs.detach() if training else probs

    # -- graph-level
    def train_step(self, graph, feat, temperature, **kwargs):
        self._check_mode(True)
        self._to(graph.device)
        labels = self.model(graph, feat, embed=False, **kwargs).argmax(-1).detach()
        prob, _ = self.explain_graph(graph, feat, temperature, training=True, **kwargs)
        return self.loss(prob, labels)

    def explain_graph(self, graph, feat, temperature=1.0, training=False, **kwargs):
        self._check_mode(True)
        self._to(graph.device)
        z = self.model(graph, feat, embed=True, **kwargs).detach()
        u, v = graph.edges()
        gates = self._gate(torch.cat([z[u.long()], z[v.long()]], dim=-1), temperature, training)
        mask = 0.5 * (gates + _reverse_values(u, v, graph.num_nodes(), gates))
        self.set_masks(graph, mask)
        probs = F.softmax(self.model(graph, feat, edge_weight=mask, **kwargs), dim=-1)
        return self._finish(probs, training), mask

    # -- node-level
    def train_step_node(self, nodes, graph, feat, temperature, **kwargs):
        self._check_mode(False)
        self._to(graph.device)
        labels = self.model(graph, feat, embed=False, **kwargs).argmax(-1).detach()
        prob, _, bg, centres = self.explain_node(nodes, graph, feat, temperature, training=True, **kwargs)
        return self.loss(prob[centres], labels[bg.ndata[NID].long()][centres])

    def explain_node(self, nodes, graph, feat, temperature=1.0, training=False, **kwargs):
        from ..batch import batch as batch_graphs
        self._check_mode(False)
        self._to(graph.device)
        nodes = [nodes] if isinstance(nodes, int) else (nodes.tolist() if torch.is_tensor(nodes) else list(nodes))
        pieces, pair_emb = [], []
        for nid in nodes:
            sg, centre = khop_in_subgraph(graph, nid, self.num_hops)
            sg.ndata["feat"] = feat[sg.ndata[NID].long()]
            marker = torch.zeros(sg.num_nodes(), dtype=torch.bool, device=sg.device)
            marker[centre.long()] = True
            sg.ndata["train"] = marker
            z = self.model(sg, sg.ndata["feat"], embed=True, **kwargs).detach()
            u, v = sg.edges()
            pair_emb.append(torch.cat([z[u.long()], z[v.long()], z[centre[0]].expand(sg.num_edges(), -1)], dim=-1))
            pieces.append(sg)
        bg = batch_graphs(pieces)
        mask = self._gate(torch.cat(pair_emb), temperature, training)
        self.set_masks(bg, mask)
        probs = F.softmax(self.model(bg, bg.ndata["feat"], edge_weight=mask, **kwargs), dim=-1)
        return self._finish(probs, training), mask, bg, bg.ndata["train"].nonzero().squeeze(1)


class HeteroPGExplainer(PGExplainer):
    """PGExplainer over typed graphs: embeddings are scored on the homogenised edge list and the gates are handed back to
    the model keyed by canonical edge type."""

    @staticmethod
    def _homogenise_with(graph, embeddings):
        from ..convert import to_homogeneous
        for nt, z in embeddings.items():
            graph.nodes[nt].data["emb"] = z.detach()
        return to_homogeneous(graph, ndata=["emb"])

    def _edge_mask_to_heterogeneous(self, edge_mask, homograph, heterograph):
        rel_of_edge = homograph.edata[ETYPE]
        return {rel: edge_mask[(rel_of_edge == heterograph.get_etype_id(rel)).nonzero().squeeze(1)]
                for rel in heterograph.canonical_etypes}

    def explain_graph(self, graph, feat, temperature=1.0, training=False, **kwargs):
        self._check_mode(True)
        self._to(graph.device)
        hg = self._homogenise_with(graph, self.model(graph, feat, embed=True, **kwargs))
        z = hg.ndata["emb"]
        u, v = hg.edges()
        gates = self._gate(torch.cat([z[u.long()], z[v.long()]], dim=-1), temperature, training)
        mask = 0.5 * (gates + _reverse_values(u, v, hg.num_nodes(), gates))
        self.set_masks(hg, mask)
        typed = self._edge_mask_to_heterogeneous(mask, hg, graph)
        probs = F.softmax(self.model(graph, feat, edge_weight=typed, **kwargs), dim=-1)
        return self._finish(probs, training), typed

    def train_step_node(self, nodes, graph, feat, temperature, **kwargs):
        self._check_mode(False)
        self._to(graph.device)
        prob, _, bg, centres = self.explain_node(nodes, graph, feat, temperature, training=True, **kwargs)
        labels = {nt: y.argmax(-1).detach() for nt, y in self.model(bg, self.batched_feats, embed=False, **kwargs).items()}
        return self.loss(torch.cat([prob[nt][ids] for nt, ids in centres.items()]),
                         torch.cat([labels[nt][ids] for nt, ids in centres.items()]))

    def explain_node(self, nodes, graph, feat, temperature=1.0, training=False, **kwargs):
        from ..batch import batch as batch_graphs
        self._check_mode(False)
        self._to(graph.device)
        homo_pieces, typed_pieces, pair_emb = [], [], []
        for nt_target, ids in nodes.items():
            ids = ids.tolist() if torch.is_tensor(ids) else ([ids] if isinstance(ids, int) else list(ids))
            for nid in ids:
                sg, centre = khop_in_subgraph(graph, {nt_target: nid}, self.num_hops)
                for nt in sg.ntypes:
                    sg.nodes[nt].data["feat"] = feat[nt][sg.nodes[nt].data[NID].long()]
                    marker = torch.zeros(sg.num_nodes(nt), dtype=torch.bool, device=sg.device)
                    if nt in centre:
                        marker[centre[nt].long()] = True
                    sg.nodes[nt].data["train"] = marker
                sg_feat = {nt: sg.nodes[nt].data["feat"] for nt in sg.ntypes}
                hg = self._homogenise_with(sg, self.model(sg, sg_feat, embed=True, **kwargs))
                z = hg.ndata["emb"]
                u, v = hg.edges()
                # position of the target node inside the homogenised id space: types are laid out in ntype order
                offset = sum(sg.num_nodes(t) for t in sg.ntypes[:sg.ntypes.index(nt_target)])
                anchor = z[offset + int(centre[nt_target][0])]
                pair_emb.append(torch.cat([z[u.long()], z[v.long()], anchor.expand(hg.num_edges(), -1)], dim=-1))
                homo_pieces.append(hg)
                typed_pieces.append(sg)
        bh, bt = batch_graphs(homo_pieces), batch_graphs(typed_pieces)
        gates = self._gate(torch.cat(pair_emb), temperature, training)
        u, v = bh.edges()
        mask = 0.5 * (gates + _reverse_values(u, v, bh.num_nodes(), gates))
        self.set_masks(bh, mask)
        typed = self._edge_mask_to_heterogeneous(mask, bh, bt)
        feats = {nt: bt.nodes[nt].data["feat"] for nt in bt.ntypes}
        logits = self.model(bt, feats, edge_weight=typed, **kwargs)
        probs = {nt: F.softmax(y, dim=-1) for nt, y in logits.items()}
        centres = {nt: bt.nodes[nt].data["train"].nonzero().squeeze(1) for nt in bt.ntypes}
        if training:
            self.batched_feats = feats
            probs = {nt: p.detach() for nt, p in probs.items()}
        else:
            self.clear_masks()
        return probs, typed, bt, centres


# --------------------------------------------------------------------------------------------------------- SubgraphX
class MCTSNode:
    """Search-tree node: a set of graph nodes plus visit statistics."""

    def __init__(self, nodes):
        self.nodes = nodes
        self.num_visit = 0
        self.total_reward = 0.0
        self.immediate_reward = 0.0
        self.children = []

    def __repr__(self):
        return str(self.nodes)


def _component_labels(num_nodes, u, v):
    """Weakly-connected component id per node: min-label propagation to a fix point (device side)."""
    label = torch.arange(num_nodes, device=u.device)
    u, v = u.long(), v.long()
    while True:
        low = torch.minimum(label[u], label[v])
        nxt = label.scatter_reduce(0, u, low, "amin").scatter_reduce(0, v, low, "amin")
        nxt = nxt[nxt]                                   # pointer jumping
        if torch.equal(nxt, label):
            return label
        label = nxt


def _largest_component(num_nodes, u, v, without=None):
    """node ids of the largest weakly-connected component (node ``without`` is not eligible)."""
    label = _component_labels(num_nodes, u, v)
    if without is not None:
        label = label.clone()
        label[without] = -1
    ids, counts = torch.unique(label[label >= 0], return_counts=True)
    if ids.numel() == 0:
        return ids
    return (label == ids[counts.argmax()]).nonzero().squeeze(1)


class SubgraphX(nn.Module):
    """Monte-Carlo tree search over connected induced subgraphs, each scored by a sampled Shapley value of the model's
    probability for ``target_class`` (``model(graph, feat)`` returns graph-level logits).

    ``coalition_batch > 1`` evaluates that many sampled coalitions per forward on a batched replica of the graph (the model
    must use batch-aware readouts, as all readouts of this package are)."""

    def __init__(self, model, num_hops, coef=10.0, high2low=True, num_child=12, num_rollouts=20, node_min=3,
                 shapley_steps=100, log=False, coalition_batch=1):
        super().__init__()
        self.model = model
        self.num_hops, self.coef, self.high2low, self.num_child = num_hops, coef, high2low, num_child
        self.num_rollouts, self.node_min, self.shapley_steps, self.log = num_rollouts, node_min, shapley_steps, log
        self.coalition_batch = max(1, int(coalition_batch))

    # -- value function
    def _neighbourhood(self, seed_mask):
        """nodes within ``num_hops - 1`` hops (either direction) of the seed set, as a boolean mask."""
        u, v = self.graph.edges()
        u, v = u.long(), v.long()
        reach = seed_mask.clone()
        for _ in range(self.num_hops - 1):
            grown = reach.clone()
            grown[v[reach[u]]] = True
            grown[u[reach[v]]] = True
            reach = grown
        return reach

    def _class_prob(self, keep_rows):
        """P(target class) for each row of ``keep_rows`` ((k, N) {0,1} node masks applied to the features)."""
        out = []
        if self.coalition_batch > 1:
            from ..batch import batch as batch_graphs
            for chunk in keep_rows.split(self.coalition_batch):
                k = chunk.shape[0]
                if k not in self._replicas:
                    self._replicas[k] = batch_graphs([self.graph] * k)
                x = (self.feat.unsqueeze(0) * chunk.unsqueeze(-1).to(self.feat.dtype)).reshape(k * self.feat.shape[0], -1)
                out.append(self.model(self._replicas[k], x, **self.kwargs).softmax(dim=-1)[:, self.target_class])
        else:
            for row in keep_rows:
                x = self.feat * row.unsqueeze(-1).to(self.feat.dtype)
                out.append(self.model(self.graph, x, **self.kwargs).softmax(dim=-1)[:, self.target_class])
        return torch.cat(out)

    def shapley(self, subgraph_nodes):
        n = self.graph.num_nodes()
        dev = self.feat.device
        inside = torch.zeros(n, dtype=torch.bool, device=dev)
        inside[torch.as_tensor(subgraph_nodes, device=dev).long()] = True
        region = self._neighbourhood(inside)
        players = (region & ~inside).nonzero().squeeze(1)                   # the other players of the local game
        steps, p = self.shapley_steps, players.numel()
        # a uniformly random permutation of players + the subgraph "player": the players ordered before it join the coalition
        rank = torch.rand(steps, p + 1, device=dev).argsort(dim=1).argsort(dim=1)
        joined = rank[:, :p] < rank[:, p:]
        without = (~region).unsqueeze(0).repeat(steps, 1)
        without[:, players] = joined
        with_sub = without | inside.unsqueeze(0)
        with torch.no_grad():
            gain = self._class_prob(with_sub) - self._class_prob(without)
        return gain.mean().item()

    # -- tree search
    def _key(self, nodes):
        return str(sorted(torch.as_tensor(nodes).tolist()))

    def get_mcts_children(self, mcts_node):
        if mcts_node.children:
            return mcts_node.children
        sub = node_subgraph(self.graph, mcts_node.nodes)
        parent_ids = sub.ndata[NID].long()
        degree = sub.in_degrees() + sub.out_degrees()
        candidates = torch.topk(degree, min(sub.num_nodes(), self.num_child), largest=self.high2low).indices
        su, sv = sub.edges()
        seen = set()
        for drop in candidates.tolist():
            if sub.num_nodes() - 1 <= 1:
                continue
            alive = (su != drop) & (sv != drop)
            keep = _largest_component(sub.num_nodes(), su[alive], sv[alive], without=drop)
            if keep.numel() == 0:
                continue
            nodes = parent_ids[keep]
            key = self._key(nodes)
            child = self.mcts_node_maps.setdefault(key, MCTSNode(nodes))
            if key not in seen:
                seen.add(key)
                mcts_node.children.append(child)
        for child in mcts_node.children:
            if child.immediate_reward == 0:
                child.immediate_reward = self.shapley(child.nodes)
        return mcts_node.children

    def mcts_rollout(self, mcts_node):
        if len(mcts_node.nodes) <= self.node_min:
            return mcts_node.immediate_reward
        children = self.get_mcts_children(mcts_node)
        if not children:
            return mcts_node.immediate_reward
        explore = self.coef * math.sqrt(sum(c.num_visit for c in children))

        def ucb(c):
            return c.total_reward / max(c.num_visit, 1) + explore * c.immediate_reward / (1 + c.num_visit)

        best = max(children, key=ucb)
        reward = self.mcts_rollout(best)
        best.num_visit += 1
        best.total_reward += reward
        return reward

    def explain_graph(self, graph, feat, target_class, **kwargs):
        self.model.eval()
        assert graph.num_nodes() > self.node_min, \
            f"The number of nodes in the graph {graph.num_nodes()} should be bigger than {self.node_min}."
        self.graph, self.feat, self.target_class, self.kwargs = graph, feat, target_class, kwargs
        self._replicas = {}
        root = MCTSNode(graph.nodes())
        self.mcts_node_maps = {self._key(root.nodes): root}
        for _ in range(self.num_rollouts):
            self.mcts_rollout(root)
        leaves = [m for m in self.mcts_node_maps.values() if len(m.nodes) <= self.node_min]
        return max(leaves, key=lambda m: m.immediate_reward).nodes if leaves else root.nodes


class HeteroSubgraphX(nn.Module):
    """SubgraphX for typed graphs. The search runs on the homogenised topology with a 1-column "presence" feature; the
    wrapped model turns presence back into per-type feature masks, so scoring uses the original typed model. Returns the
    selected node ids per node type."""

    def __init__(self, model, num_hops, coef=10.0, high2low=True, num_child=12, num_rollouts=20, node_min=3,
                 shapley_steps=100, log=False):
        super().__init__()
        self.model = model
        self._cfg = dict(num_hops=num_hops, coef=coef, high2low=high2low, num_child=num_child, num_rollouts=num_rollouts,
                         node_min=node_min, shapley_steps=shapley_steps, log=log)
        for k, v in self._cfg.items():
            setattr(self, k, v)

    def explain_graph(self, graph, feat, target_class, **kwargs):
        from ..convert import to_homogeneous
        homo = to_homogeneous(graph)
        type_of, local_id = homo.ndata[NTYPE], homo.ndata[NID]
        typed_model, ntypes = self.model, graph.ntypes
        rows = {nt: (type_of == i).nonzero().squeeze(1) for i, nt in enumerate(ntypes)}

        class _PresenceModel(nn.Module):
            def forward(self, _g, presence, **kw):
                x = {}
                for nt in feat:
                    gate = torch.zeros(graph.num_nodes(nt), 1, device=presence.device, dtype=feat[nt].dtype)
                    gate[local_id[rows[nt]].long()] = presence[rows[nt]].to(feat[nt].dtype)
                    x[nt] = feat[nt] * gate
                return typed_model(graph, x, **kw)

        search = SubgraphX(_PresenceModel(), **self._cfg)
        picked = search.explain_graph(homo, torch.ones(homo.num_nodes(), 1, device=homo.device), target_class, **kwargs).long()
        return {nt: local_id[picked[type_of[picked] == i]] for i, nt in enumerate(ntypes)}
