"""Building blocks of :class:`~dgl_b200.data.CSVDataset` (parity: /root/reference/python/dgl/data/csv_dataset_base.py).

``meta.yaml`` is validated into small dataclasses (the reference uses pydantic models of the same names), each CSV
becomes a :class:`NodeData` / :class:`EdgeData` / :class:`GraphData` table, and :class:`DGLGraphConstructor` turns
tables into graphs. User ids are arbitrary (ints or strings); the id -> row mapping is done with one sort + binary
search per table rather than a Python dict lookup per edge."""
from __future__ import annotations

import ast
import os
from dataclasses import dataclass, field, fields
from typing import Callable, List, Optional

import numpy as np
import pandas as pd
import torch
import yaml

from ..base import DGLError, dgl_warning
from ..convert import heterograph as dgl_heterograph

__all__ = ["MetaNode", "MetaEdge", "MetaGraph", "MetaYaml", "load_yaml_with_sanity_check", "BaseData", "NodeData",
           "EdgeData", "GraphData", "DGLGraphConstructor", "DefaultDataParser"]


def _build(cls, raw, where):
    """Instantiate a meta dataclass from a YAML mapping with field/type validation."""
    if not isinstance(raw, dict):
        raise DGLError(f"Validation Error for YAML fields: {where} must be a mapping, got {type(raw).__name__}.")
    known = {f.name: f for f in fields(cls)}
    unknown = set(raw) - set(known)
    if unknown:
        raise DGLError(f"Validation Error for YAML fields: unknown key(s) {sorted(unknown)} in {where}.")
    missing = [n for n, f in known.items() if f.default is field().default and f.default_factory is field().default_factory
               and n not in raw]
    if missing:
        raise DGLError(f"Validation Error for YAML fields: missing required key(s) {missing} in {where}.")
    obj = cls(**raw)
    obj.validate(where)
    return obj


def _want_str(value, name, where):
    if not isinstance(value, str):
        raise DGLError(f"Validation Error for YAML fields: {where}.{name} must be a string, got {value!r}.")


@dataclass
class MetaNode:
    """One ``node_data`` entry."""
    file_name: str
    ntype: Optional[str] = "_V"
    graph_id_field: Optional[str] = "graph_id"
    node_id_field: Optional[str] = "node_id"

    def validate(self, where):
        for n in ("file_name", "ntype", "graph_id_field", "node_id_field"):
            _want_str(getattr(self, n), n, where)


@dataclass
class MetaEdge:
    """One ``edge_data`` entry."""
    file_name: str
    etype: Optional[List[str]] = field(default_factory=lambda: ["_V", "_E", "_V"])
    graph_id_field: Optional[str] = "graph_id"
    src_id_field: Optional[str] = "src_id"
    dst_id_field: Optional[str] = "dst_id"

    def validate(self, where):
        for n in ("file_name", "graph_id_field", "src_id_field", "dst_id_field"):
            _want_str(getattr(self, n), n, where)
        if not (isinstance(self.etype, (list, tuple)) and len(self.etype) == 3 and all(isinstance(x, str) for x in self.etype)):
            raise DGLError(f"Validation Error for YAML fields: {where}.etype must be [src_type, edge_type, dst_type].")


@dataclass
class MetaGraph:
    """The optional ``graph_data`` entry."""
    file_name: str
    graph_id_field: Optional[str] = "graph_id"

    def validate(self, where):
        for n in ("file_name", "graph_id_field"):
            _want_str(getattr(self, n), n, where)


@dataclass
class MetaYaml:
    """Whole ``meta.yaml``."""
    dataset_name: str
    node_data: List[MetaNode]
    edge_data: List[MetaEdge]
    version: Optional[str] = "1.0.0"
    separator: Optional[str] = ","
    graph_data: Optional[MetaGraph] = None

    def validate(self, where):
        _want_str(self.dataset_name, "dataset_name", where)
        _want_str(self.separator, "separator", where)
        if not isinstance(self.node_data, list) or not isinstance(self.edge_data, list):
            raise DGLError("Validation Error for YAML fields: node_data and edge_data must be lists.")
        self.node_data = [_build(MetaNode, x, f"node_data[{i}]") for i, x in enumerate(self.node_data)]
        self.edge_data = [_build(MetaEdge, x, f"edge_data[{i}]") for i, x in enumerate(self.edge_data)]
        if self.graph_data is not None:
            self.graph_data = _build(MetaGraph, self.graph_data, "graph_data")


def load_yaml_with_sanity_check(yaml_file):
    """Parse + validate ``meta.yaml``: schema, supported version, unique node / edge type names."""
    with open(yaml_file) as f:
        raw = yaml.safe_load(f)
    meta = _build(MetaYaml, raw, "meta.yaml")
    if str(meta.version) != "1.0.0":
        raise DGLError(f"Invalid CSVDataset version {meta.version}. Supported versions: '1.0.0'")
    ntypes = [m.ntype for m in meta.node_data]
    if len(ntypes) > len(set(ntypes)):
        raise DGLError(f"Each node CSV file must have a unique node type name, but found duplicate node type: {ntypes}.")
    etypes = [tuple(m.etype) for m in meta.edge_data]
    if len(etypes) > len(set(etypes)):
        raise DGLError(f"Each edge CSV file must have a unique edge type name, but found duplicate edge type: {etypes}.")
    return meta


def _same_length(columns):
    sizes = {k: len(v) for k, v in columns.items()}
    if len(set(sizes.values())) > 1:
        raise DGLError(f"All data are required to have same length while some of them does not. Length of data={sizes}")


def _tensor(data):
    """numpy -> tensor; float64 is narrowed to the framework default float32."""
    t = torch.from_numpy(np.array(data, copy=True))
    return t.float() if t.dtype == torch.float64 else t


class BaseData:
    """CSV reading shared by the three table kinds."""

    @staticmethod
    def read_csv(file_name, base_dir, separator):
        path = file_name if base_dir is None else os.path.join(base_dir, file_name)
        return pd.read_csv(path, sep=separator)

    @staticmethod
    def pop_from_dataframe(df, item):
        """Remove and return a column as a 1-D array (``None`` if absent)."""
        if item not in df.columns:
            return None
        return np.atleast_1d(df.pop(item).to_numpy().squeeze())


def _group_rows(graph_id):
    """{graph id: row positions} in first-seen row order inside each group."""
    order = np.argsort(graph_id, kind="stable")
    uniq, start = np.unique(graph_id[order], return_index=True)
    return {g: rows for g, rows in zip(uniq.tolist(), np.split(order, start[1:]))}


class NodeData(BaseData):
    """Rows of one node type: user ids, per-graph membership and feature columns."""

    def __init__(self, node_id, data, type=None, graph_id=None):  # noqa: A002 - reference argument name
        self.id = np.asarray(node_id)
        self.data = data
        self.type = type if type is not None else "_V"
        self.graph_id = np.asarray(graph_id) if graph_id is not None else np.zeros(len(self.id), dtype=np.int64)
        _same_length({"id": self.id, "graph_id": self.graph_id, **self.data})

    @staticmethod
    def load_from_csv(meta: MetaNode, data_parser: Callable, base_dir=None, separator=","):
        df = BaseData.read_csv(meta.file_name, base_dir, separator)
        ids = BaseData.pop_from_dataframe(df, meta.node_id_field)
        gids = BaseData.pop_from_dataframe(df, meta.graph_id_field)
        if ids is None:
            raise DGLError(f"Missing node id field [{meta.node_id_field}] in file [{meta.file_name}].")
        return NodeData(ids, data_parser(df), type=meta.ntype, graph_id=gids)

    @staticmethod
    def to_dict(node_data: List["NodeData"]) -> dict:
        """{graph id: {ntype: {'ids': user ids in node order, 'sorted'/'rank': lookup arrays, 'data': tensors}}}."""
        out = {}
        for nd in node_data:
            for gid, rows in _group_rows(nd.graph_id).items():
                ids = nd.id[rows]
                order = np.argsort(ids, kind="stable")
                s = ids[order]
                if len(s) > 1 and np.any(s[1:] == s[:-1]):
                    raise DGLError("Node IDs are required to be unique but the following ids are duplicate: "
                                   f"{np.unique(s[1:][s[1:] == s[:-1]])}")
                out.setdefault(gid, {})[nd.type] = {
                    "ids": ids, "sorted": s, "rank": order, "dtype": ids.dtype,
                    "mapping": _LazyMapping(s, order),
                    "data": {k: _tensor(np.asarray(v)[rows]) for k, v in nd.data.items()}}
        return out


class _LazyMapping:
    """user id -> node index (dict-like, backed by the sorted id array)."""

    def __init__(self, sorted_ids, rank):
        self._s, self._r = sorted_ids, rank

    def __len__(self):
        return len(self._s)

    def __getitem__(self, key):
        p = np.searchsorted(self._s, key)
        if p >= len(self._s) or self._s[p] != key:
            raise KeyError(key)
        return int(self._r[p])

    def lookup(self, keys, what):
        keys = np.asarray(keys).astype(self._s.dtype, copy=False) if len(self._s) else np.asarray(keys)
        if len(self._s) == 0:
            if len(keys):
                raise DGLError(f"{what} ids {keys[:5]} refer to nodes that are not defined.")
            return np.empty(0, dtype=np.int64)
        p = np.minimum(np.searchsorted(self._s, keys), len(self._s) - 1)
        bad = self._s[p] != keys
        if bad.any():
            raise DGLError(f"{what} ids {np.unique(keys[bad])[:5]} do not appear in the node CSV.")
        return self._r[p].astype(np.int64)


class EdgeData(BaseData):
    """Rows of one canonical edge type."""

    def __init__(self, src_id, dst_id, data, type=None, graph_id=None):  # noqa: A002
        self.src, self.dst = np.asarray(src_id), np.asarray(dst_id)
        self.data = data
        self.type = tuple(type) if type is not None else ("_V", "_E", "_V")
        self.graph_id = np.asarray(graph_id) if graph_id is not None else np.zeros(len(self.src), dtype=np.int64)
        _same_length({"src": self.src, "dst": self.dst, "graph_id": self.graph_id, **self.data})

    @staticmethod
    def load_from_csv(meta: MetaEdge, data_parser: Callable, base_dir=None, separator=","):
        df = BaseData.read_csv(meta.file_name, base_dir, separator)
        src = BaseData.pop_from_dataframe(df, meta.src_id_field)
        if src is None:
            raise DGLError(f"Missing src id field [{meta.src_id_field}] in file [{meta.file_name}].")
        dst = BaseData.pop_from_dataframe(df, meta.dst_id_field)
        if dst is None:
            raise DGLError(f"Missing dst id field [{meta.dst_id_field}] in file [{meta.file_name}].")
        gids = BaseData.pop_from_dataframe(df, meta.graph_id_field)
        return EdgeData(src, dst, data_parser(df), type=tuple(meta.etype), graph_id=gids)

    @staticmethod
    def to_dict(edge_data: List["EdgeData"], node_dict: dict) -> dict:
        out = {}
        for ed in edge_data:
            st, _, dt = ed.type
            for gid, rows in _group_rows(ed.graph_id).items():
                if gid in out and ed.type in out[gid]:
                    raise DGLError(f"Duplicate edge type[{ed.type}] for same graph[{gid}], please place the same "
                                   "edge_type for same graph into single EdgeData.")
                nodes = node_dict.get(gid, {})
                if st not in nodes or dt not in nodes:
                    raise DGLError(f"Edge type {ed.type} of graph {gid} refers to a node type without node data.")
                u = nodes[st]["mapping"].lookup(ed.src[rows], "src")
                v = nodes[dt]["mapping"].lookup(ed.dst[rows], "dst")
                out.setdefault(gid, {})[ed.type] = {
                    "edges": (torch.from_numpy(u), torch.from_numpy(v)),
                    "data": {k: _tensor(np.asarray(val)[rows]) for k, val in ed.data.items()}}
        return out


class GraphData(BaseData):
    """Graph-level rows (labels / features of whole graphs)."""

    def __init__(self, graph_id, data):
        self.graph_id = np.asarray(graph_id)
        self.data = data
        _same_length({"graph_id": self.graph_id, **self.data})

    @staticmethod
    def load_from_csv(meta: MetaGraph, data_parser: Callable, base_dir=None, separator=","):
        df = BaseData.read_csv(meta.file_name, base_dir, separator)
        gids = BaseData.pop_from_dataframe(df, meta.graph_id_field)
        if gids is None:
            raise DGLError(f"Missing graph id field [{meta.graph_id_field}] in file [{meta.file_name}].")
        return GraphData(gids, data_parser(df))

    @staticmethod
    def to_dict(graph_data: "GraphData", graphs_dict: dict):
        """(graphs in graph-CSV order, {name: [num_graphs, -1] tensor}); ids only present in the graph CSV give
        empty graphs, ids missing from it are an error."""
        missing = np.setdiff1d(np.array(list(graphs_dict.keys())), graph_data.graph_id)
        if len(missing):
            raise DGLError(f"Found following graph ids in node/edge CSVs but not in graph CSV: {missing}.")
        graphs = []
        for gid in graph_data.graph_id.tolist():
            if gid not in graphs_dict:
                graphs_dict[gid] = dgl_heterograph({("_V", "_E", "_V"): ([], [])})
            graphs.append(graphs_dict[gid])
        data = {k: _tensor(np.asarray(v)).reshape(len(graphs), -1) for k, v in graph_data.data.items()}
        return graphs, data


class DGLGraphConstructor:
    """Tables -> graphs."""

    @staticmethod
    def construct_graphs(node_data, edge_data, graph_data=None):
        node_data = node_data if isinstance(node_data, list) else [node_data]
        edge_data = edge_data if isinstance(edge_data, list) else [edge_data]
        node_dict = NodeData.to_dict(node_data)
        edge_dict = EdgeData.to_dict(edge_data, node_dict)
        graph_dict = DGLGraphConstructor._construct_graphs(node_dict, edge_dict)
        if graph_data is None:
            graph_data = GraphData(np.zeros(1, dtype=np.int64), {})
        return GraphData.to_dict(graph_data, graph_dict)

    @staticmethod
    def _construct_graphs(node_dict, edge_dict):
        graphs = {}
        for gid, by_type in node_dict.items():
            rels = edge_dict.get(gid) or {("_V", "_E", "_V"): {"edges": ([], []), "data": {}}}
            g = dgl_heterograph({et: e["edges"] for et, e in rels.items()},
                                num_nodes_dict={nt: len(nd["ids"]) for nt, nd in by_type.items()})
            for nt, nd in by_type.items():
                for k, v in nd["data"].items():
                    g.nodes[nt].data[k] = v
            for et, e in rels.items():
                for k, v in e["data"].items():
                    g.edges[et].data[k] = v
            graphs[gid] = g
        return graphs


class DefaultDataParser:
    """DataFrame -> {column: numpy array}: header-less ("Unnamed") columns are dropped with a warning, string cells
    such as ``"0.1, 0.2"`` or ``"[1, 2]"`` are parsed into numeric rows, everything else keeps pandas' dtype."""

    def __call__(self, df: pd.DataFrame):
        out = {}
        for header in df:
            if "Unnamed" in str(header):
                dgl_warning("Unnamed column is found. Ignored...")
                continue
            col = np.atleast_1d(df[header].to_numpy().squeeze())
            if len(col) and isinstance(col[0], str):
                col = np.array([ast.literal_eval(cell) for cell in col])
            out[header] = col
        return out
