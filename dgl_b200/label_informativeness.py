"""Label informativeness (parity: /root/reference/python/dgl/label_informativeness.py)."""
from __future__ import annotations

import torch

__all__ = ["edge_label_informativeness", "node_label_informativeness"]


def check_pytorch():
    """The measures are written against torch tensors (the reference guards on its backend switch here)."""
    return True


def edge_label_informativeness(graph, y, eps=1e-8):
    y = y.long()
    num_classes = int(y.max()) + 1
    src, dst = graph.edges()
    degrees = graph.in_degrees().float()
    deg_sums = torch.zeros(num_classes, device=y.device).index_add_(0, y, degrees)
    class_degree_weighted_probs = deg_sums / deg_sums.sum()
    edge_probs = torch.zeros(num_classes, num_classes, device=y.device)
    edge_probs.index_put_((y[src.long()], y[dst.long()]), torch.ones(src.shape[0], device=y.device), accumulate=True)
    edge_probs = edge_probs / edge_probs.sum()
    edge_probs = edge_probs + eps
    numerator = (edge_probs * torch.log(edge_probs)).sum()
    denominator = (class_degree_weighted_probs * torch.log(class_degree_weighted_probs + eps)).sum()
    return (2 - numerator / denominator).item()


def node_label_informativeness(graph, y, eps=1e-8):
    y = y.long()
    num_classes = int(y.max()) + 1
    src, dst = graph.edges()
    class_probs = torch.bincount(y, minlength=num_classes).float() / y.shape[0]
    degrees = graph.in_degrees().float()
    deg_sums = torch.zeros(num_classes, device=y.device).index_add_(0, y, degrees)
    class_degree_weighted_probs = deg_sums / deg_sums.sum()
    num_nonzero = (degrees > 0).sum()
    edge_probs = torch.zeros(num_classes, num_classes, device=y.device)
    w = 1.0 / degrees[dst.long()].clamp(min=1)
    edge_probs.index_put_((y[src.long()], y[dst.long()]), w, accumulate=True)
    edge_probs = edge_probs / num_nonzero
    edge_probs = edge_probs + eps
    log = torch.log(edge_probs / (class_degree_weighted_probs[:, None] * class_probs[None, :] + eps))
    numerator = (edge_probs * log).sum()
    denominator = (class_probs * torch.log(class_probs + eps)).sum()
    return (-numerator / denominator).item()
