"""Shared synthetic problems for the example scripts (no network access: every example builds its own data)."""
import numpy as np
import torch

import dgl_b200 as dgl


def planted_partition(blocks=3, size=80, p_in=0.12, p_out=0.006, feat=16, noise=1.5, seed=0, device="cpu"):
    """Stochastic-block-model graph whose blocks are the classes, with noisy class-mean features and a 20% train mask."""
    rng = np.random.RandomState(seed)
    n = blocks * size
    y = np.repeat(np.arange(blocks), size)
    adj = np.triu(rng.rand(n, n) < np.where(y[:, None] == y[None, :], p_in, p_out), 1)
    u, v = np.nonzero(adj | adj.T)
    g = dgl.add_self_loop(dgl.graph((torch.tensor(u), torch.tensor(v)), num_nodes=n)).to(device)
    x = torch.tensor(rng.randn(blocks, feat)[y] + noise * rng.randn(n, feat), dtype=torch.float32, device=device)
    train = torch.zeros(n, dtype=torch.bool, device=device)
    train[torch.tensor(rng.permutation(n)[: n // 5])] = True
    return g, x, torch.tensor(y, device=device), train


def accuracy(logits, y, mask):
    return float((logits[mask].argmax(1) == y[mask]).float().mean())
