"""Explainability modules (parity: /root/reference/python/dgl/nn/pytorch/explain/{gnnexplainer,pgexplainer,
subgraphx}.py): GNNExplainer, HeteroGNNExplainer, PGExplainer, HeteroPGExplainer, SubgraphX, HeteroSubgraphX."""
from __future__ import annotations

import math
from math import sqrt

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..base import EID, NID
from ..subgraph import khop_in_subgraph, node_subgraph

__all__ = ["GNNExplainer", "HeteroGNNExplainer", "PGExplainer", "HeteroPGExplainer", "SubgraphX", "HeteroSubgraphX",
           "MCTSNode"]


class GNNExplainer(nn.Module):
    """Learn soft feature and edge masks maximising mutual information with the model prediction.
    The model's forward must accept (graph, feat, eweight=...)."""

    def __init__(self, model, num_hops, lr=0.01, num_epochs=100, *, alpha1=0.005, alpha2=1.0, beta1=1.0, beta2=0.1,
                 log=True):
        super().__init__()
        self.model, self.num_hops, self.lr, self.num_epochs = model, num_hops, lr, num_epochs
        self.alpha1, self.alpha2, self.beta1, self.beta2, self.log = alpha1, alpha2, beta1, beta2, log

    def _init_masks(self, graph, feat):
        num_nodes, feat_size = feat.size()
        num_edges = graph.num_edges()
        device = feat.device
        std = 0.1
        feat_mask = nn.Parameter(torch.randn(1, feat_size, device=device) * std)
        std = nn.init.calculate_gain("relu") * sqrt(2.0 / (2 * max(num_nodes, 1)))
        edge_mask = nn.Parameter(torch.randn(num_edges, device=device) * std)
        return feat_mask, edge_mask

    def _loss_regularize(self, loss, feat_mask, edge_mask):
        eps = 1e-15
        edge_mask = edge_mask.sigmoid()
        loss = loss + self.alpha1 * torch.sum(edge_mask)
        ent = -edge_mask * torch.log(edge_mask + eps) - (1 - edge_mask) * torch.log(1 - edge_mask + eps)
        loss = loss + self.alpha2 * ent.mean()
        feat_mask = feat_mask.sigmoid()
        loss = loss + self.beta1 * torch.mean(feat_mask)
        ent = -feat_mask * torch.log(feat_mask + eps) - (1 - feat_mask) * torch.log(1 - feat_mask + eps)
        loss = loss + self.beta2 * ent.mean()
        return loss

    def explain_node(self, node_id, graph, feat, **kwargs):
        self.model = self.model.to(graph.device)
        self.model.eval()
        sg, inverse_indices = khop_in_subgraph(graph, node_id, self.num_hops)
        sg_nodes = sg.ndata[NID].long()
        feat = feat[sg_nodes]
        for key, item in kwargs.items():
            if torch.is_tensor(item) and item.size(0) == graph.num_nodes():
                kwargs[key] = item[sg_nodes]
        with torch.no_grad():
            logits = self.model(graph=sg, feat=feat, **kwargs)
            pred_label = logits.argmax(dim=-1)
        feat_mask, edge_mask = self._init_masks(sg, feat)
        optimizer = torch.optim.Adam([feat_mask, edge_mask], lr=self.lr)
        for _ in range(self.num_epochs):
            optimizer.zero_grad()
            h = feat * feat_mask.sigmoid()
            logits = self.model(graph=sg, feat=h, eweight=edge_mask.sigmoid(), **kwargs)
            log_probs = logits.log_softmax(dim=-1)
            loss = -log_probs[inverse_indices, pred_label[inverse_indices]]
            loss = self._loss_regularize(loss.sum(), feat_mask, edge_mask)
            loss.backward()
            optimizer.step()
        return inverse_indices, sg, feat_mask.detach().sigmoid().squeeze(), edge_mask.detach().sigmoid()

    def explain_graph(self, graph, feat, **kwargs):
        self.model = self.model.to(graph.device)
        self.model.eval()
        with torch.no_grad():
            logits = self.model(graph=graph, feat=feat, **kwargs)
            pred_label = logits.argmax(dim=-1)
        feat_mask, edge_mask = self._init_masks(graph, feat)
        optimizer = torch.optim.Adam([feat_mask, edge_mask], lr=self.lr)
        for _ in range(self.num_epochs):
            optimizer.zero_grad()
            h = feat * feat_mask.sigmoid()
            logits = self.model(graph=graph, feat=h, eweight=edge_mask.sigmoid(), **kwargs)
            log_probs = logits.log_softmax(dim=-1)
            loss = -log_probs[0, pred_label[0]]
            loss = self._loss_regularize(loss, feat_mask, edge_mask)
            loss.backward()
            optimizer.step()
        return feat_mask.detach().sigmoid().squeeze(), edge_mask.detach().sigmoid()


class HeteroGNNExplainer(GNNExplainer):
    """GNNExplainer for heterographs: masks are dicts keyed by node type / canonical edge type."""

    def _init_masks(self, graph, feat):
        device = graph.device
        feat_masks, edge_masks = {}, {}
        std = 0.1
        for nt, f in feat.items():
            feat_masks[nt] = nn.Parameter(torch.randn(1, f.size(1), device=device) * std)
        for c in graph.canonical_etypes:
            n = graph.num_nodes(c[0]) + graph.num_nodes(c[2])
            s = nn.init.calculate_gain("relu") * sqrt(2.0 / max(n, 1))
            edge_masks[c] = nn.Parameter(torch.randn(graph.num_edges(c), device=device) * s)
        return feat_masks, edge_masks

    def _loss_regularize(self, loss, feat_masks, edge_masks):
        for m in edge_masks.values():
            loss = GNNExplainer._loss_regularize(self, loss, torch.zeros(1, 1, device=m.device), m) - \
                self.beta1 * 0.5 - self.beta2 * math.log(2)
        for m in feat_masks.values():
            fm = m.sigmoid()
            eps = 1e-15
            loss = loss + self.beta1 * fm.mean()
            ent = -fm * torch.log(fm + eps) - (1 - fm) * torch.log(1 - fm + eps)
            loss = loss + self.beta2 * ent.mean()
        return loss

    def explain_graph(self, graph, feat, **kwargs):
        self.model = self.model.to(graph.device)
        self.model.eval()
        with torch.no_grad():
            logits = self.model(graph=graph, feat=feat, **kwargs)
            pred_label = logits.argmax(dim=-1)
        feat_masks, edge_masks = self._init_masks(graph, feat)
        params = list(feat_masks.values()) + list(edge_masks.values())
        optimizer = torch.optim.Adam(params, lr=self.lr)
        for _ in range(self.num_epochs):
            optimizer.zero_grad()
            h = {nt: feat[nt] * feat_masks[nt].sigmoid() for nt in feat}
            ew = {c: edge_masks[c].sigmoid() for c in edge_masks}
            logits = self.model(graph=graph, feat=h, eweight=ew, **kwargs)
            log_probs = logits.log_softmax(dim=-1)
            loss = -log_probs[0, pred_label[0]]
            loss = self._loss_regularize(loss, feat_masks, edge_masks)
            loss.backward()
            optimizer.step()
        return ({nt: m.detach().sigmoid().squeeze() for nt, m in feat_masks.items()},
                {c: m.detach().sigmoid() for c, m in edge_masks.items()})

    def explain_node(self, ntype, node_id, graph, feat, **kwargs):
        sg, inverse_indices = khop_in_subgraph(graph, {ntype: node_id}, self.num_hops)
        inv = inverse_indices[ntype]
        sg_feat = {nt: feat[nt][sg.nodes[nt].data[NID].long()] for nt in feat}
        self.model.eval()
        with torch.no_grad():
            logits = self.model(graph=sg, feat=sg_feat, **kwargs)[ntype]
            pred_label = logits.argmax(dim=-1)
        feat_masks, edge_masks = self._init_masks(sg, sg_feat)
        optimizer = torch.optim.Adam(list(feat_masks.values()) + list(edge_masks.values()), lr=self.lr)
        for _ in range(self.num_epochs):
            optimizer.zero_grad()
            h = {nt: sg_feat[nt] * feat_masks[nt].sigmoid() for nt in sg_feat}
            ew = {c: edge_masks[c].sigmoid() for c in edge_masks}
            logits = self.model(graph=sg, feat=h, eweight=ew, **kwargs)[ntype]
            log_probs = logits.log_softmax(dim=-1)
            loss = -log_probs[inv, pred_label[inv]].sum()
            loss = self._loss_regularize(loss, feat_masks, edge_masks)
            loss.backward()
            optimizer.step()
        return (inv, sg, {nt: m.detach().sigmoid().squeeze() for nt, m in feat_masks.items()},
                {c: m.detach().sigmoid() for c, m in edge_masks.items()})


class PGExplainer(nn.Module):
    """Parameterised explainer: an MLP predicts edge masks from the embeddings of the edge's endpoints."""

    def __init__(self, model, num_features, num_hops=None, explain_graph=True, coff_budget=0.01, coff_connect=5e-4,
                 sample_bias=0.0):
        super().__init__()
        self.model = model
        self.graph_explanation = explain_graph
        self.num_features = num_features * (2 if self.graph_explanation else 3)
        self.num_hops = num_hops
        self.coff_budget, self.coff_connect, self.sample_bias = coff_budget, coff_connect, sample_bias
        self.init_bias = 0.0
        self.elayers = nn.Sequential(nn.Linear(self.num_features, 64), nn.ReLU(), nn.Linear(64, 1))

    def set_masks(self, graph, edge_mask=None):
        self.edge_mask = edge_mask

    def clear_masks(self):
        self.edge_mask = None

    def parameters(self):
        return self.elayers.parameters()

    def loss(self, prob, ori_pred):
        target_prob = prob.gather(-1, ori_pred.unsqueeze(-1))
        target_prob = target_prob + 1e-6
        pred_loss = torch.mean(-torch.log(target_prob))
        edge_mask = self.sparse_mask_values
        if self.coff_budget <= 0:
            size_loss = self.coff_budget * torch.sum(edge_mask)
        else:
            size_loss = self.coff_budget * F.relu(torch.sum(edge_mask) - self.coff_budget)
        scale = 0.99
        edge_mask = self.edge_mask * (2 * scale - 1.0) + (1.0 - scale)
        mask_ent = -edge_mask * torch.log(edge_mask) - (1 - edge_mask) * torch.log(1 - edge_mask)
        mask_ent_loss = self.coff_connect * torch.mean(mask_ent)
        return pred_loss + size_loss + mask_ent_loss

    def concrete_sample(self, w, beta=1.0, training=True):
        if training:
            bias = self.sample_bias
            random_noise = torch.rand(w.size()).to(w.device)
            random_noise = bias + (1 - 2 * bias) * random_noise
            gate_inputs = torch.log(random_noise) - torch.log(1.0 - random_noise)
            gate_inputs = (gate_inputs + w) / beta
            return torch.sigmoid(gate_inputs)
        return torch.sigmoid(w)

    def train_step(self, graph, feat, temperature, **kwargs):
        assert self.graph_explanation, '"explain_graph" must be True when initializing the module.'
        self.model = self.model.to(graph.device)
        self.elayers = self.elayers.to(graph.device)
        pred = self.model(graph, feat, embed=False, **kwargs).argmax(-1).data
        prob, _ = self.explain_graph(graph, feat, temperature, training=True, **kwargs)
        return self.loss(prob, pred)

    def train_step_node(self, nodes, graph, feat, temperature, **kwargs):
        assert not self.graph_explanation, '"explain_graph" must be False when initializing the module.'
        self.model = self.model.to(graph.device)
        self.elayers = self.elayers.to(graph.device)
        pred = self.model(graph, feat, embed=False, **kwargs).argmax(-1).data
        prob, _, batched_graph, inverse_indices = self.explain_node(nodes, graph, feat, temperature, training=True, **kwargs)
        pred = pred[batched_graph.ndata[NID].long()][inverse_indices]
        prob = prob[inverse_indices]
        return self.loss(prob, pred)

    def explain_graph(self, graph, feat, temperature=1.0, training=False, **kwargs):
        assert self.graph_explanation, '"explain_graph" must be True when initializing the module.'
        self.model = self.model.to(graph.device)
        self.elayers = self.elayers.to(graph.device)
        embed = self.model(graph, feat, embed=True, **kwargs).data
        col, row = graph.edges()
        emb = torch.cat([embed[col.long()], embed[row.long()]], dim=-1)
        emb = self.elayers(emb)
        values = emb.reshape(-1)
        values = self.concrete_sample(values, beta=temperature, training=training)
        self.sparse_mask_values = values
        # symmetrise: average the mask of (u,v) and (v,u)
        n = graph.num_nodes()
        key = col.long() * n + row.long()
        rkey = row.long() * n + col.long()
        order = torch.argsort(key)
        pos = torch.searchsorted(key[order], rkey).clamp(max=key.numel() - 1)
        has_rev = key[order][pos] == rkey
        rev_vals = torch.where(has_rev, values[order][pos], values)
        edge_mask = (values + rev_vals) / 2
        self.clear_masks()
        self.set_masks(graph, edge_mask)
        logits = self.model(graph, feat, edge_weight=self.edge_mask, **kwargs)
        probs = F.softmax(logits, dim=-1)
        if training:
            probs = probs.data
        else:
            self.clear_masks()
        return (probs, edge_mask)

    def explain_node(self, nodes, graph, feat, temperature=1.0, training=False, **kwargs):
        assert not self.graph_explanation, '"explain_graph" must be False when initializing the module.'
        assert self.num_hops is not None, '"num_hops" must be provided when initializing the module.'
        from ..batch import batch as batch_graphs
        if isinstance(nodes, int):
            nodes = [nodes]
        if torch.is_tensor(nodes):
            nodes = nodes.tolist()
        self.model = self.model.to(graph.device)
        self.elayers = self.elayers.to(graph.device)
        batched_graph, batched_embed = [], []
        for node_id in nodes:
            sg, inverse_indices = khop_in_subgraph(graph, node_id, self.num_hops)
            sg.ndata["feat"] = feat[sg.ndata[NID].long()]
            sg.ndata["train"] = torch.tensor([nid in inverse_indices for nid in range(sg.num_nodes())], device=sg.device)
            embed = self.model(sg, sg.ndata["feat"], embed=True, **kwargs).data
            col, row = sg.edges()
            self_embed = embed[inverse_indices[0]].repeat(sg.num_edges(), 1)
            emb = torch.cat([embed[col.long()], embed[row.long()], self_embed], dim=-1)
            batched_embed.append(emb)
            batched_graph.append(sg)
        batched_graph = batch_graphs(batched_graph)
        batched_embed = torch.cat(batched_embed)
        batched_embed = self.elayers(batched_embed)
        values = batched_embed.reshape(-1)
        values = self.concrete_sample(values, beta=temperature, training=training)
        self.sparse_mask_values = values
        edge_mask = values
        self.clear_masks()
        self.set_masks(batched_graph, edge_mask)
        batched_feats = batched_graph.ndata["feat"]
        logits = self.model(batched_graph, batched_feats, edge_weight=self.edge_mask, **kwargs)
        probs = F.softmax(logits, dim=-1)
        batched_inverse_indices = batched_graph.ndata["train"].nonzero().squeeze(1)
        if training:
            probs = probs.data
        else:
            self.clear_masks()
        return (probs, edge_mask, batched_graph, batched_inverse_indices)


class HeteroPGExplainer(PGExplainer):
    """PGExplainer on a heterograph: runs on the homogenised graph and maps edge masks back per etype."""

    def explain_graph(self, graph, feat, temperature=1.0, training=False, **kwargs):
        from ..convert import to_homogeneous
        self.model = self.model.to(graph.device)
        self.elayers = self.elayers.to(graph.device)
        embed = self.model(graph, feat, embed=True, **kwargs)
        for nt, emb in embed.items():
            graph.nodes[nt].data["emb"] = emb.data
        homo = to_homogeneous(graph, ndata=["emb"])
        homo_embed = homo.ndata["emb"]
        col, row = homo.edges()
        emb = torch.cat([homo_embed[col.long()], homo_embed[row.long()]], dim=-1)
        values = self.concrete_sample(self.elayers(emb).reshape(-1), beta=temperature, training=training)
        self.sparse_mask_values = values
        self.edge_mask = values
        from ..base import ETYPE
        et = homo.edata[ETYPE]
        hetero_edge_mask = {c: values[et == i] for i, c in enumerate(graph.canonical_etypes)}
        logits = self.model(graph, feat, edge_weight=hetero_edge_mask, **kwargs)
        probs = F.softmax(logits, dim=-1)
        if training:
            probs = probs.data
        else:
            self.clear_masks()
        return (probs, hetero_edge_mask)


class MCTSNode:
    def __init__(self, nodes):
        self.nodes = nodes
        self.num_visit = 0
        self.total_reward = 0.0
        self.immediate_reward = 0.0
        self.children = []

    def __repr__(self):
        return str(self.nodes)


class SubgraphX(nn.Module):
    """Monte-Carlo tree search over connected subgraphs scored by (approximate) Shapley values."""

    def __init__(self, model, num_hops, coef=10.0, high2low=True, num_child=12, num_rollouts=20, node_min=3,
                 shapley_steps=100, log=False):
        super().__init__()
        self.num_hops, self.coef, self.high2low, self.num_child = num_hops, coef, high2low, num_child
        self.num_rollouts, self.node_min, self.shapley_steps, self.log = num_rollouts, node_min, shapley_steps, log
        self.model = model

    def shapley(self, subgraph_nodes):
        num_nodes = self.graph.num_nodes()
        subgraph_nodes = subgraph_nodes.tolist()
        from ..subgraph import khop_in_subgraph as _k
        local_region = subgraph_nodes
        for _ in range(self.num_hops - 1):
            in_neighbors, _ = self.graph.in_edges(local_region)
            _, out_neighbors = self.graph.out_edges(local_region)
            local_region = list(set(local_region + in_neighbors.tolist() + out_neighbors.tolist()))
        split_point = num_nodes
        coalition_space = list(set(local_region) - set(subgraph_nodes)) + [split_point]
        marginal_contributions = []
        device = self.feat.device
        for _ in range(self.shapley_steps):
            permuted_space = np.random.permutation(coalition_space)
            split_idx = int(np.where(permuted_space == split_point)[0])
            selected_nodes = permuted_space[:split_idx]
            exclude_mask = torch.ones(num_nodes)
            exclude_mask[local_region] = 0.0
            exclude_mask[selected_nodes] = 1.0
            include_mask = exclude_mask.clone()
            include_mask[subgraph_nodes] = 1.0
            exclude_feat = self.feat * exclude_mask.unsqueeze(1).to(device)
            include_feat = self.feat * include_mask.unsqueeze(1).to(device)
            with torch.no_grad():
                exclude_probs = self.model(self.graph, exclude_feat, **self.kwargs).softmax(dim=-1)
                exclude_value = exclude_probs[:, self.target_class]
                include_probs = self.model(self.graph, include_feat, **self.kwargs).softmax(dim=-1)
                include_value = include_probs[:, self.target_class]
            marginal_contributions.append(include_value - exclude_value)
        return torch.cat(marginal_contributions).mean().item()

    def get_mcts_children(self, mcts_node):
        if len(mcts_node.children) > 0:
            return mcts_node.children
        subg = node_subgraph(self.graph, mcts_node.nodes)
        node_degrees = subg.out_degrees() + subg.in_degrees()
        k = min(subg.num_nodes(), self.num_child)
        chosen_nodes = torch.topk(node_degrees, k, largest=self.high2low).indices
        mcts_children_maps = dict()
        for node in chosen_nodes:
            new_subg = subg.clone()
            new_subg.remove_nodes(node.to(new_subg.idtype), store_ids=True)
            if new_subg.num_nodes() == 1:
                continue
            # keep the largest weakly connected component
            import scipy.sparse as sp
            from scipy.sparse.csgraph import connected_components
            u, v = new_subg.edges()
            m = sp.coo_matrix((np.ones(u.shape[0]), (u.cpu().numpy(), v.cpu().numpy())), shape=(new_subg.num_nodes(),) * 2)
            ncomp, labels = connected_components(m, directed=False)
            big = np.argmax(np.bincount(labels))
            keep = torch.from_numpy(np.nonzero(labels == big)[0])
            new_nodes = subg.ndata[NID][new_subg.ndata[NID].long()[keep]] if NID in subg.ndata else keep
            str_nodes = str(sorted(new_nodes.tolist()))
            if str_nodes not in self.mcts_node_maps:
                child_mcts_node = MCTSNode(new_nodes)
                self.mcts_node_maps[str_nodes] = child_mcts_node
            else:
                child_mcts_node = self.mcts_node_maps[str_nodes]
            if str_nodes not in mcts_children_maps:
                mcts_node.children.append(child_mcts_node)
                mcts_children_maps[str_nodes] = child_mcts_node
        for child in mcts_node.children:
            if child.immediate_reward == 0:
                child.immediate_reward = self.shapley(child.nodes)
        return mcts_node.children

    def mcts_rollout(self, mcts_node):
        if len(mcts_node.nodes) <= self.node_min:
            return mcts_node.immediate_reward
        children_nodes = self.get_mcts_children(mcts_node)
        if not children_nodes:
            return mcts_node.immediate_reward
        children_visit_sum = sum(c.num_visit for c in children_nodes)
        children_visit_sum_sqrt = math.sqrt(children_visit_sum)
        chosen_child = max(children_nodes, key=lambda c: c.total_reward / max(c.num_visit, 1)
                           + self.coef * c.immediate_reward * children_visit_sum_sqrt / (1 + c.num_visit))
        reward = self.mcts_rollout(chosen_child)
        chosen_child.num_visit += 1
        chosen_child.total_reward += reward
        return reward

    def explain_graph(self, graph, feat, target_class, **kwargs):
        self.model.eval()
        assert graph.num_nodes() > self.node_min, f"The number of nodes in the graph {graph.num_nodes()} should be bigger than {self.node_min}."
        self.graph, self.feat, self.target_class, self.kwargs = graph, feat, target_class, kwargs
        self.mcts_node_maps = dict()
        root = MCTSNode(graph.nodes())
        self.mcts_node_maps[str(root)] = root
        for _ in range(self.num_rollouts):
            self.mcts_rollout(root)
        best_leaf = None
        best_immediate_reward = float("-inf")
        for mcts_node in self.mcts_node_maps.values():
            if len(mcts_node.nodes) > self.node_min:
                continue
            if mcts_node.immediate_reward > best_immediate_reward:
                best_leaf = mcts_node
                best_immediate_reward = best_leaf.immediate_reward
        return best_leaf.nodes if best_leaf is not None else root.nodes


class HeteroSubgraphX(SubgraphX):
    """SubgraphX on heterographs via homogenisation of node ids (node sets returned per type)."""

    def explain_graph(self, graph, feat, target_class, **kwargs):
        from ..convert import to_homogeneous
        homo = to_homogeneous(graph)
        from ..base import NTYPE
        ntype_ids = homo.ndata[NTYPE]
        local = homo.ndata[NID]
        model = self.model
        ntypes = graph.ntypes

        class _Wrap(nn.Module):
            def forward(_self, g, masked_feat, **kw):
                masks = masked_feat[:, 0]
                hf = {nt: feat[nt] * masks[ntype_ids == i].unsqueeze(1) for i, nt in enumerate(ntypes)}
                return model(graph, hf, **kw)

        sub = SubgraphX(_Wrap(), self.num_hops, self.coef, self.high2low, self.num_child, self.num_rollouts,
                        self.node_min, self.shapley_steps, self.log)
        nodes = sub.explain_graph(homo, torch.ones(homo.num_nodes(), 1, device=homo.device), target_class, **kwargs)
        return {nt: local[nodes.long()][ntype_ids[nodes.long()] == i] for i, nt in enumerate(ntypes)}
