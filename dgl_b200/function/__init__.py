"""Builtin message / reduce function descriptors (``dgl.function``).

Parity: /root/reference/python/dgl/function/{message,reducer,base}.py. A builtin
is only a *descriptor*; ``core.message_passing`` lowers (message, reduce) pairs to
one fused g-SpMM and standalone messages to g-SDDMM.
"""
from __future__ import annotations

import sys
from itertools import product

__all__ = ["TargetCode", "BuiltinFunction", "MessageFunction", "ReduceFunction", "BinaryMessageFunction", "CopyMessageFunction",
           "SimpleReduceFunction", "copy_u", "copy_e", "copy_src", "copy_edge", "sum", "max", "min", "mean"]


class TargetCode:
    """Numeric codes of the three message-passing targets (reference function/base.py:7)."""
    SRC = 0
    DST = 1
    EDGE = 2
    CODE2STR = {0: "u", 1: "v", 2: "e"}


class BuiltinFunction:
    @property
    def name(self):
        raise NotImplementedError


class MessageFunction(BuiltinFunction):
    pass


class ReduceFunction(BuiltinFunction):
    pass


_TARGET_NAME = {"u": "u", "v": "v", "e": "e"}


class BinaryMessageFunction(MessageFunction):
    """``out = binary_op(lhs_target.lhs_field, rhs_target.rhs_field)``"""

    def __init__(self, binary_op, lhs, rhs, lhs_field, rhs_field, out_field):
        self.binary_op = binary_op
        self.lhs = lhs
        self.rhs = rhs
        self.lhs_field = lhs_field
        self.rhs_field = rhs_field
        self.out_field = out_field

    @property
    def name(self):
        return f"{self.lhs}_{self.binary_op}_{self.rhs}"

    def __repr__(self):
        return f"fn.{self.name}('{self.lhs_field}', '{self.rhs_field}', '{self.out_field}')"


class CopyMessageFunction(MessageFunction):
    def __init__(self, target, in_field, out_field):
        self.target = target
        self.in_field = in_field
        self.out_field = out_field

    @property
    def name(self):
        return f"copy_{self.target}"

    def __repr__(self):
        return f"fn.{self.name}('{self.in_field}', '{self.out_field}')"


class SimpleReduceFunction(ReduceFunction):
    def __init__(self, name, msg_field, out_field):
        self._name = name
        self.msg_field = msg_field
        self.out_field = out_field

    @property
    def name(self):
        return self._name

    def __repr__(self):
        return f"fn.{self._name}('{self.msg_field}', '{self.out_field}')"


def copy_u(u, out):
    """Message = source node feature ``u``."""
    return CopyMessageFunction("u", u, out)


def copy_e(e, out):
    """Message = edge feature ``e``."""
    return CopyMessageFunction("e", e, out)


def copy_src(src, out):
    return copy_u(src, out)


def copy_edge(edge, out):
    return copy_e(edge, out)


def _make_binary(lhs, op, rhs):
    def fn(lhs_field, rhs_field, out):
        return BinaryMessageFunction(op, lhs, rhs, lhs_field, rhs_field, out)

    fn.__name__ = f"{lhs}_{op}_{rhs}"
    fn.__doc__ = f"Message = {lhs}.lhs_field {op} {rhs}.rhs_field"
    return fn


_mod = sys.modules[__name__]
for _l, _r in product("uve", "uve"):
    if _l == _r:
        continue
    for _op in ("add", "sub", "mul", "div", "dot"):
        _f = _make_binary(_l, _op, _r)
        setattr(_mod, _f.__name__, _f)
        __all__.append(_f.__name__)


def _make_reduce(name):
    def fn(msg, out):
        return SimpleReduceFunction(name, msg, out)

    fn.__name__ = name
    fn.__doc__ = f"Reduce messages by {name}."
    return fn


sum = _make_reduce("sum")  # noqa: A001
max = _make_reduce("max")  # noqa: A001
min = _make_reduce("min")  # noqa: A001
mean = _make_reduce("mean")

# legacy aliases (reference: function/message.py src_mul_edge etc.)
def src_mul_edge(src, edge, out):
    return getattr(_mod, "u_mul_e")(src, edge, out)


def __getattr__(name):
    raise AttributeError(f"module 'dgl_b200.function' has no attribute '{name}'")
