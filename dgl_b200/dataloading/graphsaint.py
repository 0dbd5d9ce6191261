"""GraphSAINT subgraph samplers (Zeng et al., ICLR 2020; role of the reference's dataloading/graphsaint.py).

Each call draws a node set and returns the subgraph it induces; the three published strategies differ only in how the
node set is drawn:

* ``'node'``  — ``budget`` node draws with probability proportional to out-degree;
* ``'edge'``  — ``budget`` edge draws with probability proportional to 1/deg_in(dst) + 1/deg_out(src), keeping both
  end points;
* ``'walk'``  — ``budget = (num_roots, walk_length)``: uniform roots, one random walk each (native walk kernel), all
  visited nodes.

The sampling distribution depends only on the graph, so it is computed once and reused when ``cache`` is on."""
from __future__ import annotations

import torch

from ..base import DGLError
from ..sampling import random_walk
from .base import Sampler, set_edge_lazy_features, set_node_lazy_features

_MODES = ("node", "edge", "walk")


class SAINTSampler(Sampler):
    def __init__(self, mode, budget, cache=True, prefetch_ndata=None, prefetch_edata=None, output_device="cpu"):
        super().__init__()
        if mode not in _MODES:
            raise DGLError(f"Expect mode to be 'node', 'edge' or 'walk', got {mode}.")
        self.mode, self.budget, self.cache = mode, budget, cache
        self.prob = None                                  # cached sampling distribution (node or edge weights)
        self.prefetch_ndata = list(prefetch_ndata or [])
        self.prefetch_edata = list(prefetch_edata or [])
        self.output_device = output_device
        self.sampler = getattr(self, mode + "_sampler")   # kept as an attribute: callers may swap in their own

    # ---- distributions ------------------------------------------------------------------------------------------
    def _weights(self, g):
        """Unnormalised draw weights for the current mode (cached per sampler when ``cache``)."""
        if self.cache and self.prob is not None:
            return self.prob
        if self.mode == "node":
            w = g.out_degrees().clamp(min=1).float()
        else:
            u, v = g.edges()
            inv_in = g.in_degrees().clamp(min=1).float().reciprocal()
            inv_out = g.out_degrees().clamp(min=1).float().reciprocal()
            w = inv_in[v.long()] + inv_out[u.long()]
            w = w / w.sum()
        if self.cache:
            self.prob = w
        return w

    def _draw(self, weights, count):
        return torch.unique(torch.multinomial(weights, num_samples=int(count), replacement=True))

    # ---- node-set strategies ---------------------------------------------------------------------------------------
    def node_sampler(self, g):
        return self._draw(self._weights(g), self.budget).to(g.idtype)

    def edge_sampler(self, g):
        picked = self._draw(self._weights(g), self.budget)
        u, v = g.edges()
        return torch.unique(torch.cat([u[picked], v[picked]])).to(g.idtype)

    def walk_sampler(self, g):
        num_roots, walk_length = self.budget
        roots = torch.randint(0, g.num_nodes(), (int(num_roots),), device=g.device)
        traces, _ = random_walk(g, nodes=roots, length=int(walk_length))
        return torch.unique(traces[traces >= 0]).to(g.idtype)      # -1 marks walks that hit a dead end

    def sample(self, g, indices):
        """``indices`` only counts iterations (one subgraph per item of the loader's index set)."""
        sg = g.subgraph(self.sampler(g), relabel_nodes=True, output_device=self.output_device)
        set_node_lazy_features(sg, self.prefetch_ndata)
        set_edge_lazy_features(sg, self.prefetch_edata)
        return sg
