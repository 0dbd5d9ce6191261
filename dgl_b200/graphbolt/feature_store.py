"""Abstract feature store: ``(domain, type, name)``-keyed collection of ``Feature`` objects.

Parity: /root/reference/python/dgl/graphbolt/feature_store.py (FeatureKey, FeatureStore); the concrete stores are in
impl/features.py (BasicFeatureStore, TorchBasedFeatureStore)."""
from __future__ import annotations

from collections import namedtuple

from .impl.features import Feature

__all__ = ["FeatureKey", "FeatureStore", "Feature"]


FeatureKey = namedtuple("FeatureKey", ["domain", "type", "name"])


class FeatureStore:
    """Abstract collection of features addressed by (domain, type, name) (reference feature_store.py)."""

    def __getitem__(self, feature_key) -> Feature:
        raise NotImplementedError

    def __setitem__(self, feature_key, feature):
        raise NotImplementedError

    def __contains__(self, feature_key) -> bool:
        raise NotImplementedError

    def read(self, domain, type_name, feature_name, ids=None):
        return self[(domain, type_name, feature_name)].read(ids)

    def size(self, domain, type_name, feature_name):
        return self[(domain, type_name, feature_name)].size()

    def count(self, domain, type_name, feature_name):
        return self[(domain, type_name, feature_name)].count()

    def metadata(self, domain, type_name, feature_name):
        return self[(domain, type_name, feature_name)].metadata()

    def update(self, domain, type_name, feature_name, value, ids=None):
        self[(domain, type_name, feature_name)].update(value, ids)

    def keys(self):
        raise NotImplementedError


def __getattr__(name):
    # names the reference keeps in this module but that live elsewhere in this package
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    for modname in ('dgl_b200.graphbolt.compat',):
        mod = importlib.import_module(modname)
        if name in vars(mod):
            return vars(mod)[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
