"""CSVDataset (parity: /root/reference/python/dgl/data/csv_dataset.py): a directory holding ``meta.yaml`` plus
nodes_*.csv / edges_*.csv and an optional graphs.csv becomes one or many (hetero)graphs.

Parsing is split in ``csv_dataset_base``: YAML validation, one table object per CSV, graph construction. Feature
columns go through a *data parser* — a callable ``DataFrame -> {name: ndarray}`` — which can be given per node type /
edge type (dict) or as one callable; :class:`~.csv_dataset_base.DefaultDataParser` is the default. Processed graphs
are cached next to the CSVs (``<name>.bin``) and reloaded unless ``force_reload``."""
from __future__ import annotations

import os

import torch

from ..base import DGLError
from .dgl_dataset import DGLDataset
from .graph_serialize import load_graphs, save_graphs
from .utils import Subset


class CSVDataset(DGLDataset):
    META_YAML_NAME = "meta.yaml"

    def __init__(self, data_path, force_reload=False, verbose=True, ndata_parser=None, edata_parser=None,
                 gdata_parser=None, transform=None):
        from .csv_dataset_base import DefaultDataParser, load_yaml_with_sanity_check
        self.graphs, self.data = None, None
        self.ndata_parser = {} if ndata_parser is None else ndata_parser
        self.edata_parser = {} if edata_parser is None else edata_parser
        self.gdata_parser = gdata_parser
        self.default_data_parser = DefaultDataParser()
        self.data_path = data_path
        meta_path = os.path.join(data_path, self.META_YAML_NAME)
        if not os.path.exists(meta_path):
            raise DGLError(f"'{self.META_YAML_NAME}' cannot be found under {data_path}.")
        self.meta_yaml = load_yaml_with_sanity_check(meta_path)
        super().__init__(self.meta_yaml.dataset_name, raw_dir=os.path.dirname(meta_path), force_reload=force_reload,
                         verbose=verbose, transform=transform)

    # the CSVs live in the directory itself, not under <raw_dir>/<name>
    @property
    def raw_path(self):
        return self.data_path

    @property
    def save_path(self):
        return self.data_path

    def _parser(self, table, key):
        if callable(table):
            return table
        return table.get(key, self.default_data_parser)

    def process(self):
        from .csv_dataset_base import DGLGraphConstructor, EdgeData, GraphData, NodeData
        meta, base, sep = self.meta_yaml, self.data_path, self.meta_yaml.separator
        nodes = [NodeData.load_from_csv(m, self._parser(self.ndata_parser, m.ntype), base_dir=base, separator=sep)
                 for m in meta.node_data]
        edges = [EdgeData.load_from_csv(m, self._parser(self.edata_parser, tuple(m.etype)), base_dir=base, separator=sep)
                 for m in meta.edge_data]
        graph_data = None
        if meta.graph_data is not None:
            graph_data = GraphData.load_from_csv(meta.graph_data, self.gdata_parser or self.default_data_parser,
                                                 base_dir=base, separator=sep)
        self.graphs, self.data = DGLGraphConstructor.construct_graphs(nodes, edges, graph_data)
        self._set_labels()

    def _set_labels(self):
        if len(self.data) == 1:
            self.labels = next(iter(self.data.values()))

    def _cache_file(self):
        return os.path.join(self.save_path, self.name + ".bin")

    def has_cache(self):
        return os.path.exists(self._cache_file())

    def save(self):
        if self.graphs is None:
            raise DGLError("No graphs available in dataset")
        save_graphs(self._cache_file(), self.graphs, labels=self.data)

    def load(self):
        self.graphs, self.data = load_graphs(self._cache_file())
        self._set_labels()

    def __getitem__(self, i):
        if torch.is_tensor(i) and i.dim() == 1:
            return Subset(self, i.cpu())
        g = self.graphs[i] if self._transform is None else self._transform(self.graphs[i])
        if len(self.data) == 1:
            return g, self.labels[i]
        if len(self.data) > 0:
            return g, {k: v[i] for k, v in self.data.items()}
        return g

    def __len__(self):
        return len(self.graphs)
