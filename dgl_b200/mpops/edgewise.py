"""Edge-wise message-passing operators with an optional edge type (parity: /root/reference/python/dgl/mpops/edgewise.py):
``copy_u``/``copy_v`` fetch endpoint data onto edges; ``u_{add,sub,mul,div,dot}_v`` combine both endpoints with
broadcasting. All of them lower to the g-SDDMM kernels through ``dgl_b200.ops``."""
import sys

from .. import ops

__all__ = ["copy_u", "copy_v"]


def _relation(g, etype):
    return g if etype is None else g[etype]


def copy_u(g, x_node, etype=None):
    """``Y_e = X_u`` for every edge e = (u, v): ``x_node`` [|V_src|, *] -> [|E|, *]."""
    return ops.gsddmm(_relation(g, etype), "copy_lhs", x_node, None)


def copy_v(g, x_node, etype=None):
    """``Y_e = X_v`` for every edge e = (u, v): ``x_node`` [|V_dst|, *] -> [|E|, *]."""
    return ops.gsddmm(_relation(g, etype), "copy_rhs", None, x_node)


def _binary(op):
    def func(g, x_node, y_node, etype=None):
        return ops.gsddmm(_relation(g, etype), op, x_node, y_node, lhs_target="u", rhs_target="v")
    func.__name__ = f"u_{op}_v"
    func.__doc__ = (f"``Z_e = {op}(X_u, Y_v)`` for every edge e = (u, v) with NumPy broadcasting over the feature "
                    f"dimensions; ``etype`` selects a relation of a heterograph.")
    return func


for _op in ("add", "sub", "mul", "div", "dot"):
    _f = _binary(_op)
    setattr(sys.modules[__name__], _f.__name__, _f)
    __all__.append(_f.__name__)
