"""Autograd for the generalized sparse ops.

The differentiation algebra is the reference's
(/root/reference/python/dgl/backend/pytorch/sparse.py:162-248 GSpMM,
:443-503 GSDDMM, :685-747 EdgeSoftmax, :828-866 SegmentReduce): the backward of
a g-SpMM is a g-SpMM on the transposed relation plus a g-SDDMM. Every
forward/backward leg calls one native kernel (sm_100a on CUDA tensors).
``gidx`` below is a :class:`dgl_b200.graph_index.UnitGraph`.
"""
from __future__ import annotations

import torch

from .. import _sparse_ops as K
from ..base import DGLError


def _reduce_grad(grad, shape):
    """Sum ``grad`` over broadcast dims so it matches ``shape``
    (reference: backend/pytorch/sparse.py:43-74)."""
    grad_shape = grad.shape[1:]
    in_shape = tuple(shape[1:])
    if tuple(grad_shape) == in_shape:
        return grad
    num_to_squeeze = len(grad_shape) - len(in_shape)
    in_shape = (1,) * num_to_squeeze + in_shape
    reduce_idx = [i + 1 for i, (a, b) in enumerate(zip(grad_shape, in_shape)) if a != b]
    if reduce_idx:
        grad = grad.sum(dim=tuple(reduce_idx), keepdim=True)
    return grad.view(-1, *shape[1:])


def _need_reduce_last_dim(ufeat, efeat):
    """reference: sparse.py:77-85 — the 'use dot instead of mul' shortcut."""
    if ufeat is None or efeat is None:
        return False
    ushp, eshp = ufeat.shape, efeat.shape
    return ushp[1:-1] == eshp[1:-1] and eshp[-1] == 1 and ushp[-1] > 1


def _expand(x, shape):
    return x.expand(-1, *shape)


def _scatter_rows(n, shape_tail, arg, vals, dtype, device):
    """dX[arg[v,f], f] += vals[v,f] for arg >= 0 (max/min backward)."""
    out = torch.zeros((n,) + tuple(shape_tail), dtype=dtype, device=device)
    flat_out = out.view(n, -1)
    a = arg.reshape(arg.shape[0], -1).long()
    v = vals.reshape(vals.shape[0], -1)
    mask = a >= 0
    a = a.clamp(min=0)
    v = torch.where(mask, v, torch.zeros((), dtype=v.dtype, device=v.device))
    flat_out.scatter_add_(0, a, v.to(dtype))
    return out


class GSpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gidx, op, reduce_op, X, Y):
        out, argX, argY = K.gspmm(gidx.csc(), op, reduce_op, X, Y, gidx.num_dst)
        ctx.backward_cache = (gidx, op, reduce_op, None if X is None else X.shape, None if Y is None else Y.shape,
                              out.dtype, out.device)
        req_x = X is not None and X.requires_grad
        req_y = Y is not None and Y.requires_grad
        # save only what backward needs (reference: sparse.py:92-128)
        save_x = X if (req_y and op in ("mul", "add", "dot") and op == "mul") else None
        save_y = Y if (req_x and op == "mul") else None
        if reduce_op in ("max", "min"):
            if op == "mul":
                save_x, save_y = (X if req_y else None), (Y if req_x else None)
        ctx.save_for_backward(save_x, save_y, argX, argY)
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, dZ):
        gidx, op, reduce_op, xshape, yshape, dtype, device = ctx.backward_cache
        X, Y, argX, argY = ctx.saved_tensors
        dX = dY = None
        dZ = dZ.contiguous()
        if op != "copy_rhs" and ctx.needs_input_grad[3]:
            if reduce_op == "sum":
                g_rev = gidx.reverse()
                if op == "mul":
                    dX, _, _ = K.gspmm(g_rev.csc(), "mul", "sum", dZ, Y, g_rev.num_dst)
                else:  # add / copy_lhs
                    dX, _, _ = K.gspmm(g_rev.csc(), "copy_lhs", "sum", dZ, None, g_rev.num_dst)
            else:
                vals = dZ
                if op == "mul":
                    Yb = _expand(Y, dZ.shape[1:]) if Y.shape[1:] != dZ.shape[1:] else Y
                    vals = dZ * Yb.reshape(Y.shape[0], -1).gather(
                        0, argY.reshape(argY.shape[0], -1).long().clamp(min=0)).view_as(dZ)
                dX = _scatter_rows(xshape[0], dZ.shape[1:], argX, vals, dtype, device)
            dX = _reduce_grad(dX, xshape)
        if op != "copy_lhs" and ctx.needs_input_grad[4]:
            if reduce_op == "sum":
                if op == "mul" and _need_reduce_last_dim(X, Y):
                    dY = K.gsddmm(gidx.coo(), "dot", X, dZ, "u", "v")
                elif op == "mul":
                    dY = K.gsddmm(gidx.coo(), "mul", X, dZ, "u", "v")
                else:  # add / copy_rhs
                    dY = K.gsddmm(gidx.coo(), "copy_rhs", None, dZ, "u", "v")
            else:
                vals = dZ
                if op == "mul":
                    Xb = _expand(X, dZ.shape[1:]) if X.shape[1:] != dZ.shape[1:] else X
                    vals = dZ * Xb.reshape(X.shape[0], -1).gather(
                        0, argX.reshape(argX.shape[0], -1).long().clamp(min=0)).view_as(dZ)
                dY = _scatter_rows(yshape[0], dZ.shape[1:], argY, vals, dtype, device)
            dY = _reduce_grad(dY, yshape)
        return None, None, None, dX, dY


class GSDDMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gidx, op, X, Y, lhs_target, rhs_target):
        out = K.gsddmm(gidx.coo(), op, X, Y, lhs_target, rhs_target, gidx=gidx)
        ctx.backward_cache = (gidx, op, lhs_target, rhs_target, None if X is None else X.shape,
                              None if Y is None else Y.shape)
        req_x = X is not None and X.requires_grad
        req_y = Y is not None and Y.requires_grad
        save_x = X if (req_y and op in ("mul", "dot", "div")) else None
        save_y = Y if ((req_x or req_y) and op in ("mul", "dot", "div")) else None
        ctx.save_for_backward(save_x, save_y)
        return out

    @staticmethod
    def backward(ctx, dZ):
        gidx, op, lt, rt, xshape, yshape = ctx.backward_cache
        X, Y = ctx.saved_tensors
        dZ = dZ.contiguous()
        dX = dY = None

        def to_target(c, target):
            # reduce a per-edge tensor c to `target` rows
            if target == "e":
                return c
            if target == "u":
                g_rev = gidx.reverse()
                return K.gspmm(g_rev.csc(), "copy_rhs", "sum", None, c, g_rev.num_dst)[0]
            return K.gspmm(gidx.csc(), "copy_rhs", "sum", None, c, gidx.num_dst)[0]

        def fused(other, other_target, target):
            # sum_e other[other_target(e)] * dZ[e] reduced to `target`; uses one g-SpMM when
            # `other` lives on the opposite endpoint
            if target == "u" and other_target == "v":
                g_rev = gidx.reverse()
                return K.gspmm(g_rev.csc(), "mul", "sum", other, dZ, g_rev.num_dst)[0]
            if target == "v" and other_target == "u":
                return K.gspmm(gidx.csc(), "mul", "sum", other, dZ, gidx.num_dst)[0]
            c = K.gsddmm(gidx.coo(), "mul", other, dZ, other_target, "e")
            return to_target(c, target)

        if op != "copy_rhs" and ctx.needs_input_grad[2]:
            if op in ("add", "sub", "copy_lhs"):
                dX = to_target(dZ, lt)
            elif op in ("mul", "dot"):
                dX = fused(Y, rt, lt)
            elif op == "div":
                c = K.gsddmm(gidx.coo(), "div", dZ, Y, "e", rt)
                dX = to_target(c, lt)
            dX = _reduce_grad(dX, xshape)
        if op != "copy_lhs" and ctx.needs_input_grad[3]:
            if op in ("add", "copy_rhs"):
                dY = to_target(dZ, rt)
            elif op == "sub":
                dY = to_target(-dZ, rt)
            elif op in ("mul", "dot"):
                dY = fused(X, lt, rt)
            elif op == "div":
                # d(x/y)/dy = -x / y^2
                c = K.gsddmm(gidx.coo(), "mul", dZ, X, "e", lt)
                y2 = Y * Y
                c = -K.gsddmm(gidx.coo(), "div", c, y2, "e", rt)
                dY = to_target(c, rt)
            dY = _reduce_grad(dY, yshape)
        return None, None, dX, dY, None, None


class EdgeSoftmax(torch.autograd.Function):
    """Fused single-kernel edge softmax (the reference composes 5 kernels on GPU:
    backend/pytorch/sparse.py:709-713)."""

    @staticmethod
    def forward(ctx, gidx, score, eids, norm_by):
        if eids is not None:
            raise DGLError("edge_softmax on an edge subset is handled by the caller")
        g = gidx if norm_by == "dst" else gidx.reverse()
        out = K.edge_softmax_forward(g.csc(), score.contiguous(), g.num_dst)
        ctx.backward_cache = g
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.backward_cache
        (out,) = ctx.saved_tensors
        gs = K.edge_softmax_backward(g.csc(), out, grad_out.contiguous(), g.num_dst)
        return None, gs, None, None


class SegmentReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op, x, offsets):
        red = "sum" if op == "mean" else op
        out, arg = K.segment_reduce(offsets, x.contiguous(), red)
        if op == "mean":
            cnt = (offsets[1:] - offsets[:-1]).clamp(min=1).to(out.dtype)
            out = out / cnt.view(-1, *([1] * (out.dim() - 1)))
        ctx.backward_cache = (op, x.shape)
        ctx.save_for_backward(arg, offsets)
        return out

    @staticmethod
    def backward(ctx, dy):
        op, xshape = ctx.backward_cache
        arg, offsets = ctx.saved_tensors
        seglen = (offsets[1:] - offsets[:-1]).long()
        if op in ("sum", "mean"):
            if op == "mean":
                dy = dy / seglen.clamp(min=1).to(dy.dtype).view(-1, *([1] * (dy.dim() - 1)))
            dx = torch.repeat_interleave(dy, seglen, dim=0, output_size=xshape[0])
        else:
            dx = _scatter_rows(xshape[0], dy.shape[1:], arg, dy, dy.dtype, dy.device)
        return None, dx, None


def _autocast_args(*tensors):
    """Under ``torch.autocast`` cast the floating CUDA operands of a sparse op to the autocast dtype (the reference does
    the same: backend/pytorch/sparse.py:146-159 ``_cast_if_autocast_enabled``), so a bf16 activation meeting an fp32
    edge weight does not raise a dtype mismatch. Returns (cast tensors, context that disables autocast inside the op)."""
    if not (torch.is_autocast_enabled() and any(t is not None and torch.is_tensor(t) and t.is_cuda for t in tensors)):
        return tensors, _NullCtx()
    dt = torch.get_autocast_dtype("cuda") if hasattr(torch, "get_autocast_dtype") else torch.get_autocast_gpu_dtype()
    cast = tuple(t.to(dt) if (t is not None and torch.is_tensor(t) and t.is_cuda and t.is_floating_point()
                              and t.dtype != torch.float64) else t for t in tensors)
    return cast, torch.autocast("cuda", enabled=False)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def gspmm(gidx, op, reduce_op, lhs_data, rhs_data):
    if op == "sub":
        op, rhs_data = "add", -rhs_data
    elif op == "div":
        op, rhs_data = "mul", 1.0 / rhs_data
    (lhs_data, rhs_data), ctx = _autocast_args(lhs_data, rhs_data)
    with ctx:
        return GSpMM.apply(gidx, op, reduce_op, lhs_data, rhs_data)


def gsddmm(gidx, op, lhs_data, rhs_data, lhs_target="u", rhs_target="v"):
    (lhs_data, rhs_data), ctx = _autocast_args(lhs_data, rhs_data)
    with ctx:
        return GSDDMM.apply(gidx, op, lhs_data, rhs_data, lhs_target, rhs_target)


def edge_softmax(gidx, logits, eids=None, norm_by="dst"):
    (logits,), ctx = _autocast_args(logits)
    with ctx:
        return EdgeSoftmax.apply(gidx, logits, eids, norm_by)


def segment_reduce(op, x, offsets):
    (x,), ctx = _autocast_args(x)
    with ctx:
        return SegmentReduce.apply(op, x, offsets)
