"""nn utilities: graph-aware ``Sequential``, ``WeightBasis``, ``JumpingKnowledge``, ``LabelPropagation`` and the
graph-construction modules ``KNNGraph`` / ``SegmentedKNNGraph`` / ``RadiusGraph``.

API parity: /root/reference/python/dgl/nn/pytorch/utils.py and factory.py. The propagation of ``LabelPropagation`` runs on
the framework's g-SpMM directly (``nn/_mp.py``), k steps without touching node frames.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..transforms import knn_graph, radius_graph, segmented_knn_graph
from . import _mp


def _is_index_vector(t):
    return t.dtype == torch.int64 and t.dim() == 1


def matmul_maybe_select(A, B):
    """``A @ B`` where a 1-D int64 ``A`` stands for a one-hot matrix (row lookup)."""
    return B[A] if _is_index_vector(A) else A @ B


def bmm_maybe_select(A, B, index):
    """Row i uses matrix ``B[index[i]]``: ``A[i] @ B[index[i]]``, or the row ``B[index[i], A[i]]`` for one-hot ``A``."""
    if _is_index_vector(A):
        return B[index, A]
    return torch.einsum("ni,nio->no", A, B[index])


class Identity(nn.Module):
    """Returns its input (a placeholder for an optional sub-module)."""

    def forward(self, x):
        return x


def _as_tuple(x):
    return x if isinstance(x, tuple) else (x,)


class Sequential(nn.Sequential):
    """Chains graph layers: every module is called as ``module(graph, *features)`` and its output feeds the next one.
    With a list of graphs, module i runs on graph i (one block per layer)."""

    def forward(self, graph, *feats):
        graphs = graph if isinstance(graph, list) else [graph] * len(self)
        for g_i, layer in zip(graphs, self):
            feats = layer(g_i, *_as_tuple(feats))
        return feats


class WeightBasis(nn.Module):
    """``num_outputs`` tensors of ``shape`` generated as learned combinations of ``num_bases`` shared bases."""

    def __init__(self, shape, num_bases, num_outputs):
        super().__init__()
        self.shape, self.num_bases, self.num_outputs = tuple(shape), num_bases, num_outputs
        if num_outputs <= num_bases:
            import warnings
            warnings.warn("The number of weight outputs should be larger than the number of bases.")
        gain = nn.init.calculate_gain("relu")
        self.weight = nn.Parameter(nn.init.xavier_uniform_(torch.empty(num_bases, *self.shape), gain=gain))
        self.w_comp = nn.Parameter(nn.init.xavier_uniform_(torch.empty(num_outputs, num_bases), gain=gain))

    def forward(self):
        mixed = self.w_comp @ self.weight.flatten(1)
        return mixed.view(self.num_outputs, *self.shape)


class JumpingKnowledge(nn.Module):
    """Combine the per-layer representations of a node: concatenation, element-wise max, or LSTM attention."""

    MODES = ("cat", "max", "lstm")

    def __init__(self, mode="cat", in_feats=None, num_layers=None):
        super().__init__()
        if mode not in self.MODES:
            raise AssertionError(f"Expect mode to be 'cat', or 'max' or 'lstm', got {mode}")
        self.mode = mode
        if mode == "lstm":
            if in_feats is None or num_layers is None:
                raise AssertionError("in_feats and num_layers are required for lstm mode")
            width = (num_layers * in_feats) // 2
            self.lstm = nn.LSTM(in_feats, width, bidirectional=True, batch_first=True)
            self.att = nn.Linear(2 * width, 1)

    def reset_parameters(self):
        for m in self.children():
            m.reset_parameters()

    def forward(self, feat_list):
        if self.mode == "cat":
            return torch.cat(feat_list, dim=-1)
        layers = torch.stack(feat_list, dim=1)                       # [N, L, D]
        if self.mode == "max":
            return layers.amax(dim=1)
        scores = self.att(self.lstm(layers)[0]).squeeze(-1)          # [N, L]
        return torch.einsum("nl,nld->nd", scores.softmax(dim=-1), layers)


class LabelPropagation(nn.Module):
    """``Y <- alpha * Â Y + (1 - alpha) * Y0`` for ``k`` steps; ``Â`` symmetric or row normalised."""

    def __init__(self, k, alpha, norm_type="sym", clamp=True, normalize=False, reset=False):
        super().__init__()
        self.k, self.alpha, self.norm_type = k, alpha, norm_type
        self.clamp, self.normalize, self.reset = clamp, normalize, reset

    def forward(self, g, labels, mask=None):
        if self.norm_type not in ("sym", "row"):
            raise ValueError(f"Expect norm_type to be 'sym' or 'row', got {self.norm_type}")
        if labels.dtype == torch.long:
            labels = F.one_hot(labels.view(-1)).float()
        known = labels if mask is None else torch.zeros_like(labels).index_put_((mask,), labels[mask])
        anchor = (1.0 - self.alpha) * known
        y = known
        mode = "both" if self.norm_type == "sym" else "right"
        for _ in range(self.k):
            y = anchor + self.alpha * _mp.sym_norm_propagate(g, y, mode)
            if self.clamp:
                y = y.clamp(0.0, 1.0)
            if self.normalize:
                y = F.normalize(y, p=1)
            if self.reset:
                y = y.clone()
                y[mask] = labels[mask]
        return y


class KNNGraph(nn.Module):
    """Points -> k-nearest-neighbour graph (``transforms.knn_graph``)."""

    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, x, algorithm="bruteforce-blas", dist="euclidean", exclude_self=False):
        return knn_graph(x, self.k, algorithm=algorithm, dist=dist, exclude_self=exclude_self)


class SegmentedKNNGraph(nn.Module):
    """k-NN graph built independently inside each segment of a batched point set."""

    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, x, segs, algorithm="bruteforce-blas", dist="euclidean", exclude_self=False):
        return segmented_knn_graph(x, self.k, segs, algorithm=algorithm, dist=dist, exclude_self=exclude_self)


class RadiusGraph(nn.Module):
    """Points -> graph connecting pairs closer than ``r`` in the ``p``-norm."""

    def __init__(self, r, p=2, self_loop=False, compute_mode="donot_use_mm_for_euclid_dist"):
        super().__init__()
        self.r, self.p, self.self_loop, self.compute_mode = r, p, self_loop, compute_mode

    def forward(self, x, get_distances=False):
        return radius_graph(x, self.r, self.p, self.self_loop, self.compute_mode, get_distances)
