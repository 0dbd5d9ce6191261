"""Typed views of the sections of an OnDiskDataset ``metadata.yaml`` (parity:
/root/reference/python/dgl/graphbolt/impl/ondisk_metadata.py, pydantic models there; plain dict subclasses with attribute
access here, so a parsed YAML mapping can be used directly)."""
from __future__ import annotations

import enum

__all__ = ["OnDiskFeatureDataFormat", "OnDiskTVTSetData", "OnDiskTVTSet", "OnDiskFeatureDataDomain", "OnDiskFeatureData",
           "OnDiskGraphTopologyType", "OnDiskGraphTopology", "OnDiskTaskData", "ExtraMetaData", "OnDiskMetaData"]


class OnDiskFeatureDataFormat(str, enum.Enum):
    TORCH = "torch"
    NUMPY = "numpy"


class OnDiskTVTSetData(dict):
    """One entry of a train / validation / test set: name, format, in_memory, path (reference ondisk_metadata.py)."""

    def __init__(self, name=None, format="numpy", in_memory=True, path=None, **kw):
        super().__init__(name=name, format=str(getattr(format, "value", format)), in_memory=in_memory, path=path, **kw)
        self.__dict__ = self


class OnDiskTVTSet(dict):
    def __init__(self, type=None, data=None, **kw):
        super().__init__(type=type, data=[OnDiskTVTSetData(**d) if not isinstance(d, OnDiskTVTSetData) else d for d in (data or [])], **kw)
        self.__dict__ = self


class OnDiskFeatureDataDomain(str, enum.Enum):
    NODE = "node"
    EDGE = "edge"
    GRAPH = "graph"


class OnDiskFeatureData(dict):
    def __init__(self, domain=None, type=None, name=None, format="numpy", path=None, in_memory=True, **kw):
        super().__init__(domain=str(getattr(domain, "value", domain)), type=type, name=name,
                         format=str(getattr(format, "value", format)), path=path, in_memory=in_memory, **kw)
        self.__dict__ = self


class OnDiskGraphTopologyType(str, enum.Enum):
    FUSED_CSC_SAMPLING = "FusedCSCSamplingGraph"


class OnDiskGraphTopology(dict):
    def __init__(self, type=OnDiskGraphTopologyType.FUSED_CSC_SAMPLING, path=None, **kw):
        super().__init__(type=str(getattr(type, "value", type)), path=path, **kw)
        self.__dict__ = self


class OnDiskTaskData(dict):
    def __init__(self, train_set=None, validation_set=None, test_set=None, **extra_fields):
        super().__init__(train_set=train_set or [], validation_set=validation_set or [], test_set=test_set or [], **extra_fields)
        self.__dict__ = self


class ExtraMetaData(dict):
    """Base of the metadata models: unknown keys are kept and grouped under ``extra_fields`` (reference
    impl/ondisk_metadata.py:24)."""

    _known = ()

    def __init__(self, **kw):
        extra = dict(kw.pop("extra_fields", None) or {})
        for k in list(kw):
            if self._known and k not in self._known:
                extra[k] = kw.pop(k)
        super().__init__(extra_fields=extra, **kw)
        self.__dict__ = self


class OnDiskMetaData(dict):
    """Top-level model of ``metadata.yaml`` (dataset_name, graph_topology, feature_data, tasks, include_original_edge_id)."""

    def __init__(self, dataset_name=None, graph_topology=None, feature_data=None, tasks=None, include_original_edge_id=False,
                 **kw):
        super().__init__(dataset_name=dataset_name, graph_topology=graph_topology, feature_data=feature_data or [],
                         tasks=tasks or [], include_original_edge_id=include_original_edge_id, **kw)
        self.__dict__ = self
