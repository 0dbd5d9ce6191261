"""Feature initializers (parity: /root/reference/python/dgl/init.py).

An initializer fills the rows that appear when nodes/edges are added or when a feature is set on a subset of rows:
``fn(shape, dtype, ctx, id_range) -> tensor``. Install with ``DGLGraph.set_n_initializer`` / ``set_e_initializer``."""
from __future__ import annotations

import torch

__all__ = ["base_initializer", "zero_initializer"]


def base_initializer(shape, dtype, ctx, id_range):  # noqa: ARG001 - this is the documented signature
    """Signature every initializer follows: ``shape`` of the block to create (rows first), its ``dtype`` and device
    ``ctx``, and ``id_range`` (a ``slice`` of the node/edge ids being initialised; step is always ``None``)."""
    raise NotImplementedError


def zero_initializer(shape, dtype, ctx, id_range):  # noqa: ARG001
    """All-zero rows (the default of every frame)."""
    return torch.zeros(shape, dtype=dtype, device=ctx)
