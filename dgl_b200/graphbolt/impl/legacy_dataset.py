"""GraphBolt view of a legacy ``dgl.data`` dataset (parity: /root/reference/python/dgl/graphbolt/impl/legacy_dataset.py)."""
from __future__ import annotations

from ..dataset import Dataset, Task

__all__ = ["LegacyDataset"]


class LegacyDataset(Dataset):
    """GraphBolt view of a legacy ``dgl.data`` node-classification dataset (reference impl/legacy_dataset.py)."""

    def __init__(self, legacy):
        from .features import BasicFeatureStore, TorchBasedFeature
        from .fused_csc_sampling_graph import from_dglgraph
        from ..itemset import ItemSet
        g = legacy[0]
        self._name = getattr(legacy, "name", type(legacy).__name__)
        self._graph = from_dglgraph(g)
        self._feature = BasicFeatureStore({("node", None, k): TorchBasedFeature(v) for k, v in g.ndata.items()
                                           if v.is_floating_point()})
        task = Task()
        task.metadata = {"name": "node_classification", "num_classes": getattr(legacy, "num_classes", None)}
        lab = g.ndata.get("label")
        for split, mask in (("train_set", "train_mask"), ("validation_set", "val_mask"), ("test_set", "test_mask")):
            if mask in g.ndata and lab is not None:
                ids = g.ndata[mask].nonzero().squeeze(1)
                setattr(task, split, ItemSet((ids, lab[ids]), names=("seeds", "labels")))
        self._tasks = [task]
        self._all = ItemSet(g.num_nodes(), names="seeds")

    tasks = property(lambda self: self._tasks)
    graph = property(lambda self: self._graph)
    feature = property(lambda self: self._feature)
    dataset_name = property(lambda self: self._name)
    all_nodes_set = property(lambda self: self._all)
