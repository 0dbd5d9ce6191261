"""Constants, error types and small helpers shared by the whole package.

Parity: /root/reference/python/dgl/base.py (ALL, NID/EID/NTYPE/ETYPE field names,
DGLError, dgl_warning).
"""
from __future__ import annotations

import warnings

import torch

__all__ = ["ALL", "NID", "EID", "NTYPE", "ETYPE", "DGLError", "DGLWarning", "dgl_warning", "dgl_warning_format", "is_all"]


class DGLError(Exception):
    """Framework error."""


class _All:
    def __repr__(self):
        return "ALL"

    def __reduce__(self):
        return "ALL"


ALL = _All()
NID = "_ID"
EID = "_ID"
NTYPE = "_TYPE"
ETYPE = "_TYPE"
_INTERNAL_COLUMNS = {NID, EID, NTYPE, ETYPE}


def is_all(x) -> bool:
    return isinstance(x, _All) or (isinstance(x, slice) and x == slice(None))


def is_internal_column(name):
    return name in _INTERNAL_COLUMNS


class DGLWarning(UserWarning):
    pass


_default_formatwarning = warnings.formatwarning


def dgl_warning_format(message, category, filename, lineno, line=None):
    """``warnings.formatwarning`` replacement: DGL warnings print as one short line, everything else as usual."""
    if isinstance(category, type) and issubclass(category, DGLWarning):
        return f"DGL Warning: {message}\n"
    return _default_formatwarning(message, category, filename, lineno, line)


def dgl_warning(message, category=DGLWarning, stacklevel=2):
    warnings.warn(message, category=category, stacklevel=stacklevel)


warnings.formatwarning = dgl_warning_format


_DTYPE_CODE = {torch.float32: 0, torch.float64: 1, torch.float16: 2, torch.bfloat16: 3}
_ID_CODE = {torch.int32: 0, torch.int64: 1}


def dtype_code(dt):
    try:
        return _DTYPE_CODE[dt]
    except KeyError:
        raise DGLError(f"unsupported feature dtype {dt}; expected float32/float64/float16/bfloat16")


def id_code(dt):
    try:
        return _ID_CODE[dt]
    except KeyError:
        raise DGLError(f"unsupported id dtype {dt}; expected int32/int64")
