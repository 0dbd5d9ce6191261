"""nn utilities (parity: /root/reference/python/dgl/nn/pytorch/utils.py, factory.py): Sequential, WeightBasis,
JumpingKnowledge, LabelPropagation, LaplacianPosEnc, Identity; KNNGraph / SegmentedKNNGraph / RadiusGraph."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import function as fn
from ..transforms import knn_graph, radius_graph, segmented_knn_graph


def matmul_maybe_select(A, B):
    if A.dtype == torch.int64 and len(A.shape) == 1:
        return B.index_select(0, A)
    return torch.matmul(A, B)


def bmm_maybe_select(A, B, index):
    if A.dtype == torch.int64 and len(A.shape) == 1:
        B = B.view(-1, B.shape[2])
        flatidx = index * B.shape[1] // B.shape[1] * 0 + index * (B.shape[0] // max(int(index.max()) + 1, 1)) + A
        return B.index_select(0, flatidx)
    BB = B.index_select(0, index)
    return torch.bmm(A.unsqueeze(1), BB).squeeze(1)


class Identity(nn.Module):
    def forward(self, x):
        return x


class Sequential(nn.Sequential):
    """Sequential container whose modules take (graph, feat...) (reference utils.py:Sequential)."""

    def forward(self, graph, *feats):
        if isinstance(graph, list):
            for graph_i, module in zip(graph, self):
                if not isinstance(feats, tuple):
                    feats = (feats,)
                feats = module(graph_i, *feats)
        else:
            for module in self:
                if not isinstance(feats, tuple):
                    feats = (feats,)
                feats = module(graph, *feats)
        return feats


class WeightBasis(nn.Module):
    """W_o = sum_b w_{o,b} V_b (reference utils.py:WeightBasis)."""

    def __init__(self, shape, num_bases, num_outputs):
        super().__init__()
        self.shape, self.num_bases, self.num_outputs = shape, num_bases, num_outputs
        if num_outputs <= num_bases:
            import warnings
            warnings.warn("The number of weight outputs should be larger than the number of bases.")
        self.weight = nn.Parameter(torch.Tensor(self.num_bases, *shape))
        nn.init.xavier_uniform_(self.weight, gain=nn.init.calculate_gain("relu"))
        self.w_comp = nn.Parameter(torch.Tensor(self.num_outputs, self.num_bases))
        nn.init.xavier_uniform_(self.w_comp, gain=nn.init.calculate_gain("relu"))

    def forward(self):
        weight = torch.matmul(self.w_comp, self.weight.view(self.num_bases, -1))
        return weight.view(self.num_outputs, *self.shape)


class JumpingKnowledge(nn.Module):
    def __init__(self, mode="cat", in_feats=None, num_layers=None):
        super().__init__()
        assert mode in ["cat", "max", "lstm"], "Expect mode to be 'cat', or 'max' or 'lstm', got {}".format(mode)
        self.mode = mode
        if mode == "lstm":
            assert in_feats is not None, "in_feats is required for lstm mode"
            assert num_layers is not None, "num_layers is required for lstm mode"
            hidden_size = (num_layers * in_feats) // 2
            self.lstm = nn.LSTM(in_feats, hidden_size, bidirectional=True, batch_first=True)
            self.att = nn.Linear(2 * hidden_size, 1)

    def reset_parameters(self):
        if self.mode == "lstm":
            self.lstm.reset_parameters()
            self.att.reset_parameters()

    def forward(self, feat_list):
        if self.mode == "cat":
            return torch.cat(feat_list, dim=-1)
        if self.mode == "max":
            return torch.stack(feat_list, dim=-1).max(dim=-1)[0]
        stacked_feat_list = torch.stack(feat_list, dim=1)
        alpha, _ = self.lstm(stacked_feat_list)
        alpha = self.att(alpha).squeeze(-1)
        alpha = torch.softmax(alpha, dim=-1)
        return (stacked_feat_list * alpha.unsqueeze(-1)).sum(dim=1)


class LabelPropagation(nn.Module):
    def __init__(self, k, alpha, norm_type="sym", clamp=True, normalize=False, reset=False):
        super().__init__()
        self.k, self.alpha, self.norm_type = k, alpha, norm_type
        self.clamp, self.normalize, self.reset = clamp, normalize, reset

    def forward(self, g, labels, mask=None):
        with g.local_scope():
            if labels.dtype == torch.long:
                labels = F.one_hot(labels.view(-1)).to(torch.float32)
            y = labels
            if mask is not None:
                y = torch.zeros_like(labels)
                y[mask] = labels[mask]
            init = (1 - self.alpha) * y
            in_degs = g.in_degrees().float().clamp(min=1)
            out_degs = g.out_degrees().float().clamp(min=1)
            if self.norm_type == "sym":
                norm_i = torch.pow(in_degs, -0.5).to(labels.device).unsqueeze(1)
                norm_j = torch.pow(out_degs, -0.5).to(labels.device).unsqueeze(1)
            elif self.norm_type == "row":
                norm_i = torch.pow(in_degs, -1.0).to(labels.device).unsqueeze(1)
            else:
                raise ValueError(f"Expect norm_type to be 'sym' or 'row', got {self.norm_type}")
            for _ in range(self.k):
                g.ndata["h"] = y * norm_j if self.norm_type == "sym" else y
                g.update_all(fn.copy_u("h", "m"), fn.sum("m", "h"))
                y = init + self.alpha * g.ndata["h"] * norm_i
                if self.clamp:
                    y = y.clamp_(0.0, 1.0)
                if self.normalize:
                    y = F.normalize(y, p=1)
                if self.reset:
                    y[mask] = labels[mask]
            return y


class KNNGraph(nn.Module):
    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, x, algorithm="bruteforce-blas", dist="euclidean", exclude_self=False):
        return knn_graph(x, self.k, algorithm=algorithm, dist=dist, exclude_self=exclude_self)


class SegmentedKNNGraph(nn.Module):
    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, x, segs, algorithm="bruteforce-blas", dist="euclidean", exclude_self=False):
        return segmented_knn_graph(x, self.k, segs, algorithm=algorithm, dist=dist, exclude_self=exclude_self)


class RadiusGraph(nn.Module):
    def __init__(self, r, p=2, self_loop=False, compute_mode="donot_use_mm_for_euclid_dist"):
        super().__init__()
        self.r, self.p, self.self_loop, self.compute_mode = r, p, self_loop, compute_mode

    def forward(self, x, get_distances=False):
        return radius_graph(x, self.r, self.p, self.self_loop, self.compute_mode, get_distances)
