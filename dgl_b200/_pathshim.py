"""Reference module paths without a file of their own.

The reference spreads one package over many small modules (``dgl.graphbolt.feature_fetcher``,
``dgl.nn.pytorch.conv.gatconv``, ``dgl.optim.pytorch.sparse_optim`` ...); this package groups the same names differently.
``from dgl_b200.<reference path> import Name`` keeps working through ONE import hook: :class:`AliasFinder` sits at the end
of ``sys.meta_path`` (so real modules always win), recognises the paths listed in ``_alias_table.ALIASES`` and builds a
module whose attributes resolve lazily (module ``__getattr__``) in the modules that actually define them - no stub files,
no eager imports, no import cycles."""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys
import types

_PKG = __name__.rsplit(".", 1)[0]


def _table():
    from ._alias_table import ALIASES, SOURCE_GROUPS
    return ALIASES, SOURCE_GROUPS


class _GuardedPackage(types.ModuleType):
    """Package type that keeps a function when an alias module of the same name is imported under it.

    ``dgl.ops.edge_softmax`` is both a function of ``dgl.ops`` and a module path of the reference; there the package's
    ``from .edge_softmax import *`` makes the function win. Here the alias module is created on demand, and the import
    system would bind it over the function afterwards - this ``__setattr__`` declines exactly that one rebinding (the
    module stays importable through ``sys.modules``)."""

    def __setattr__(self, name, value):
        held = self.__dict__.get(name)
        if getattr(value, "__alias_module__", False) and held is not None and not isinstance(held, types.ModuleType):
            return
        super().__setattr__(name, value)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, sources, names):
        self.sources, self.names = list(sources), list(names)

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        sources, names, modname = self.sources, self.names, module.__name__

        def __getattr__(name):
            if name.startswith("__"):
                raise AttributeError(name)
            for src in sources:
                try:
                    return getattr(importlib.import_module(src), name)
                except AttributeError:
                    continue
            raise AttributeError(f"module {modname!r} has no attribute {name!r}")

        module.__alias_module__ = True
        parent_name, _, leaf = modname.rpartition(".")
        parent = sys.modules.get(parent_name)
        if (parent is not None and type(parent) is types.ModuleType and leaf in parent.__dict__
                and not isinstance(parent.__dict__[leaf], types.ModuleType)):
            parent.__class__ = _GuardedPackage
        module.__getattr__ = __getattr__
        module.__dir__ = lambda: list(names)
        module.__all__ = list(names)
        module.__doc__ = (f"Reference module path ``{modname.replace(_PKG, 'dgl', 1)}``: its names are implemented in "
                          f"{', '.join(sources[:3])} ... and resolved lazily.")


class AliasFinder(importlib.abc.MetaPathFinder):
    """Resolves ``dgl_b200.<alias>`` for the aliases of ``_alias_table`` when no real module of that name exists."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PKG + "."):
            return None
        aliases, groups = _table()
        key = fullname[len(_PKG) + 1:]
        entry = aliases.get(key)
        if entry is None:
            return None
        group, names = entry
        nested = any(k.startswith(key + ".") for k in aliases)     # e.g. backend.pytorch holds backend.pytorch.sparse
        return importlib.machinery.ModuleSpec(fullname, _AliasLoader(groups[group], names), is_package=nested)


def install():
    if not any(isinstance(f, AliasFinder) for f in sys.meta_path):
        sys.meta_path.append(AliasFinder())


def alias_modules():
    """Dotted names of every synthesised module (for tests and documentation)."""
    return sorted(f"{_PKG}.{k}" for k in _table()[0])
