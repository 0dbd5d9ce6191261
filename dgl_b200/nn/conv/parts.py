"""Building blocks that the reference exposes next to its conv layers (parity:
/root/reference/python/dgl/nn/pytorch/conv/{twirlsconv,grouprevres,cugraph_base}.py).

TWIRLS pieces follow the paper's equations (Yang et al. 2021, "Graph Neural Networks Inspired by Classical Iterative
Algorithms"): the graph carries edge weights in ``edata['w']`` and weighted degrees in ``ndata['deg']``;
:class:`Attention` (IRLS re-weighting) refreshes both, :class:`Propagate` (eq. 28, pre-conditioned) and
:class:`PropagateNoPrecond` (eq. 30) take one unfolded gradient step of the energy."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import function as fn

__all__ = ["TWIRLSUnfoldingAndAttention", "AX", "normalized_AX", "D_power_X", "D_power_bias_X", "Propagate", "PropagateNoPrecond", "Attention", "MLP",
           "InvertibleCheckpoint", "CuGraphBaseConv"]


def AX(graph, X):
    """Weighted aggregation ``A @ X`` with ``A[v, u] = w(u -> v)`` taken from ``edata['w']``."""
    graph.srcdata["h"] = X
    graph.update_all(fn.u_mul_e("h", "w", "m"), fn.sum("m", "h"))
    return graph.dstdata["h"]


def D_power_X(graph, X, power):
    """``D^power @ X`` with the weighted degrees of ``ndata['deg']``."""
    return X * graph.ndata["deg"].pow(power).view(-1, 1)


def D_power_bias_X(graph, X, power, coeff, bias):
    """``(coeff * D + bias * I)^power @ X`` — the pre-conditioner of eq. 28 uses coeff = lam, bias = 1 - lam."""
    return X * (coeff * graph.ndata["deg"] + bias).pow(power).view(-1, 1)


def normalized_AX(graph, X):
    """``D^{-1/2} A D^{-1/2} X``."""
    return D_power_X(graph, AX(graph, D_power_X(graph, X, -0.5)), -0.5)


class Propagate(nn.Module):
    """One pre-conditioned, re-parameterised step (eq. 28):
    ``Z <- (1 - a) Z + a * lam * D~^{-1/2} A D~^{-1/2} Z + a * D~^{-1} Z0`` with ``D~ = lam * D + (1 - lam) I``."""

    def _prop(self, graph, Y, lam):
        Y = D_power_bias_X(graph, Y, -0.5, lam, 1 - lam)
        return D_power_bias_X(graph, AX(graph, Y), -0.5, lam, 1 - lam)

    def forward(self, graph, Y, X, alp, lam):
        return (1 - alp) * Y + alp * lam * self._prop(graph, Y, lam) + alp * D_power_bias_X(graph, X, -1, lam, 1 - lam)


class PropagateNoPrecond(nn.Module):
    """One plain gradient step with the normalised Laplacian (eq. 30):
    ``Y <- (1 - a*lam - a) Y + a * lam * D^{-1/2} A D^{-1/2} Y + a * Y0``."""

    def forward(self, graph, Y, X, alp, lam):
        return (1 - alp * lam - alp) * Y + alp * lam * normalized_AX(graph, Y) + alp * X


class Attention(nn.Module):
    """IRLS edge re-weighting: ``w_uv = 1 / max(tau, ||y_u - y_v||^{2 - p})``, edges longer than ``T`` (when T > 0) are
    cut (weight -> 0); ``ndata['deg']`` is refreshed with the new weighted degrees and dropout is applied to the
    weights afterwards so degrees are unchanged in expectation."""

    def __init__(self, tau, T, p, attn_dropout=0.0):
        super().__init__()
        self.tau, self.T, self.p, self.attn_dropout = tau, T, p, attn_dropout

    def reweighting(self, graph):
        w = F.relu(graph.edata["w"]) + 1e-7               # squared distances; clamp rounding noise below zero
        w = w.pow(1 - 0.5 * self.p).clamp(min=self.tau)
        if self.T > 0:
            w = torch.where(w > self.T, torch.full_like(w, float("inf")), w)
        graph.edata["w"] = 1.0 / w + 1e-9                 # keep isolated-by-truncation nodes at non-zero degree

    def forward(self, graph, Y, etas=None):
        if etas is not None:
            Y = Y * etas.view(-1)
        graph.srcdata["h"] = Y
        graph.apply_edges(fn.u_sub_v("h", "h", "_diff"))
        graph.edata["w"] = graph.edata.pop("_diff").square().sum(-1)
        self.reweighting(graph)
        graph.update_all(fn.copy_e("w", "m"), fn.sum("m", "deg"))
        graph.ndata["deg"] = graph.ndata["deg"].view(-1)
        if self.attn_dropout > 0:
            graph.edata["w"] = F.dropout(graph.edata["w"], self.attn_dropout, training=self.training)
        return graph


class MLP(nn.Module):
    """``num_layers`` linear layers with norm ('batch' | 'layer' | 'none') -> ReLU -> dropout between them;
    ``init_activate`` applies that block to the input first; ``num_layers == 0`` is the identity."""

    def __init__(self, input_d, hidden_d, output_d, num_layers, dropout, norm, init_activate):
        super().__init__()
        self.init_activate, self.norm, self.dropout = init_activate, norm, dropout
        dims = [input_d] + [hidden_d] * (num_layers - 1) + [output_d] if num_layers > 0 else []
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        n_norm = max(num_layers - 1, 0) + int(init_activate)
        make = {"batch": nn.BatchNorm1d, "layer": nn.LayerNorm}.get(norm)
        self.norms = nn.ModuleList([make(hidden_d) for _ in range(n_norm)]) if make is not None else None
        self.reset_params()

    def reset_params(self):
        for layer in self.layers:
            nn.init.xavier_normal_(layer.weight)
            nn.init.zeros_(layer.bias)

    def activate(self, x, k):
        if self.norms is not None:
            x = self.norms[k](x)
        return F.dropout(F.relu(x), self.dropout, training=self.training)

    def forward(self, x):
        k = 0
        if self.init_activate:
            x, k = self.activate(x, k), k + 1
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i != len(self.layers) - 1:
                x, k = self.activate(x, k), k + 1
        return x


class InvertibleCheckpoint(torch.autograd.Function):
    """Memory-free activations for an invertible block: forward stores nothing but the *output* (inputs' storage is
    released); backward first reconstructs the inputs with ``fn_inverse``, then re-runs ``fn`` under grad to obtain the
    gradients. Call as ``InvertibleCheckpoint.apply(fn, fn_inverse, num_inputs, *inputs_and_weights)`` — the first
    ``num_inputs`` tensors are the block inputs, the rest the parameters ``fn`` closes over (so they receive grads)."""

    @staticmethod
    def forward(ctx, fn_forward, fn_inverse, num_inputs, *inputs_and_weights):
        ctx.fn, ctx.fn_inverse, ctx.num_inputs = fn_forward, fn_inverse, num_inputs
        inputs = inputs_and_weights[:num_inputs]
        ctx.weights = inputs_and_weights[num_inputs:]
        ctx.input_requires_grad = [t.requires_grad for t in inputs]
        with torch.no_grad():
            out = fn_forward(*[t.detach() for t in inputs])
        single = torch.is_tensor(out)
        outs = (out,) if single else tuple(out)
        ctx.single = single
        # free the (first) input's memory: it will be rebuilt from the output in backward
        x = inputs[0]
        if x.is_leaf is False or not x.requires_grad:
            x.untyped_storage().resize_(0)
        ctx.inputs = inputs
        ctx.outputs = tuple(o.detach() for o in outs)
        return out if single else outs

    @staticmethod
    def backward(ctx, *grad_outputs):
        inputs, outs = ctx.inputs, ctx.outputs
        with torch.no_grad():
            rebuilt = ctx.fn_inverse(*outs, *[t for t in inputs[1:]])
            rebuilt = rebuilt if torch.is_tensor(rebuilt) else rebuilt[0]
            x = inputs[0]
            if x.untyped_storage().size() == 0:
                x.untyped_storage().resize_(rebuilt.untyped_storage().size())
                x.copy_(rebuilt)
        detached = [t.detach().requires_grad_(rg) for t, rg in zip(inputs, ctx.input_requires_grad)]
        with torch.enable_grad():
            re_out = ctx.fn(*detached)
        re_outs = (re_out,) if torch.is_tensor(re_out) else tuple(re_out)
        wanted = [t for t in detached if t.requires_grad] + [w for w in ctx.weights if w.requires_grad]
        grads = torch.autograd.grad(re_outs, wanted, grad_outputs, allow_unused=True) if wanted else ()
        it = iter(grads)
        g_in = [next(it) if t.requires_grad else None for t in detached]
        g_w = [next(it) if w.requires_grad else None for w in ctx.weights]
        return (None, None, None) + tuple(g_in) + tuple(g_w)


class CuGraphBaseConv(nn.Module):
    """Abstract base of the cugraph-ops backed layers (reference conv/cugraph_base.py). The CuGraph* layers of this
    package run on the native sm_100a kernels and do not need pylibcugraphops; the class keeps the shared helper."""

    def __init__(self):
        super().__init__()
        self._cached_offsets_fg = None

    def reset_parameters(self):
        raise NotImplementedError

    def forward(self, *args):
        raise NotImplementedError

    def pad_offsets(self, offsets, size):
        """CSC ``indptr`` of a block extended to ``size`` entries by repeating the last offset, i.e. the block viewed as
        a square graph whose extra destination nodes have no in-edges. The buffer is cached and grown on demand."""
        buf = self._cached_offsets_fg
        if buf is None or buf.numel() < size or buf.device != offsets.device or buf.dtype != offsets.dtype:
            buf = self._cached_offsets_fg = torch.empty(size, dtype=offsets.dtype, device=offsets.device)
        buf[: offsets.numel()] = offsets
        buf[offsets.numel(): size] = offsets[-1]
        return buf[:size]


class TWIRLSUnfoldingAndAttention(nn.Module):
    """The unfolded optimiser of TWIRLS: ``prop_step`` descent steps on the graph-regularised energy (pre-conditioned
    eq. 28 or plain eq. 30), with one IRLS attention re-weighting before the first step (``init_att``) and / or after step
    ``attn_aft``. Edge weights and weighted degrees travel in ``edata['w']`` / ``ndata['deg']`` of a private copy."""

    def __init__(self, d, alp, lam, prop_step, attn_aft=-1, tau=0.2, T=-1, p=1, use_eta=False, init_att=False,
                 attn_dropout=0, precond=True):
        super().__init__()
        self.d, self.lam, self.tau, self.p = d, lam, tau, p
        self.alp = alp if alp > 0 else 1 / (lam + 1)          # default step size
        self.prop_step, self.attn_aft, self.use_eta, self.init_att = prop_step, attn_aft, use_eta, init_att
        step = Propagate if precond else PropagateNoPrecond
        self.prop_layers = nn.ModuleList([step() for _ in range(prop_step)])
        self.init_attn = Attention(tau, T, p, attn_dropout) if init_att else None
        self.attn_layer = Attention(tau, T, p, attn_dropout) if attn_aft >= 0 else None
        self.etas = nn.Parameter(torch.ones(d)) if use_eta else None

    def forward(self, g, X):
        g = g.local_var()
        g.edata["w"] = torch.ones(g.num_edges(), 1, device=X.device, dtype=X.dtype)
        g.ndata["deg"] = g.in_degrees().to(X)
        Y = X
        if self.init_att:
            g = self.init_attn(g, Y, self.etas)
        for k, step in enumerate(self.prop_layers):
            Y = step(g, Y, X, self.alp, self.lam)
            if k == self.attn_aft - 1:
                g = self.attn_layer(g, Y, self.etas)
        return Y
