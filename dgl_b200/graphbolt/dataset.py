"""Dataset / Task protocols (parity: /root/reference/python/dgl/graphbolt/dataset.py); implementations: impl/datasets.py
(OnDiskDataset, BuiltinDataset) and impl/legacy_dataset.py."""
from __future__ import annotations

__all__ = ["Task", "Dataset"]


class Task:
    """Abstract task: metadata + train / validation / test item sets (reference dataset.py)."""

    metadata: Dict[str, Any] = None
    train_set = validation_set = test_set = None


class Dataset:
    """Abstract dataset: tasks, graph, feature store, name, all-nodes set (reference dataset.py)."""

    @property
    def tasks(self):
        raise NotImplementedError

    @property
    def graph(self):
        raise NotImplementedError

    @property
    def feature(self):
        raise NotImplementedError

    @property
    def dataset_name(self):
        raise NotImplementedError

    @property
    def all_nodes_set(self):
        raise NotImplementedError
