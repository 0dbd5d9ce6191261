"""Lazy re-export helper for reference module paths.

The reference spreads one package over many files (``dgl.graphbolt.dataloader``, ``dgl.optim.pytorch.sparse_optim``
...); this package groups the same names differently. A shim module created with :func:`reexport` makes
``from dgl_b200.<reference path> import Name`` work by resolving ``Name`` in the module(s) that define it here, on
first access (module ``__getattr__``), so no import cycles and no eager loading are introduced."""
from __future__ import annotations

import importlib


def reexport(shim_name, sources, names):
    """Return ``(__getattr__, __dir__, __all__)`` for the shim module ``shim_name``: ``names`` are looked up in each
    module of ``sources`` (dotted paths) in order."""
    names = list(names)

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        for src in sources:
            mod = importlib.import_module(src)
            try:
                return getattr(mod, name)
            except AttributeError:
                continue
        raise AttributeError(f"module {shim_name!r} has no attribute {name!r}")

    def __dir__():
        return names

    return __getattr__, __dir__, names
