"""CSVDataset (parity: /root/reference/python/dgl/data/csv_dataset.py): a directory with ``meta.yaml`` +
nodes_*.csv / edges_*.csv (+ optional graphs.csv)."""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch
import yaml

from ..base import DGLError
from ..convert import heterograph
from .dgl_dataset import DGLDataset


def _parse_feats(df, skip):
    out = {}
    for col in df.columns:
        if col in skip:
            continue
        vals = df[col].to_numpy()
        if vals.dtype == object:
            arr = np.stack([np.array([float(x) for x in str(v).split(",")]) for v in vals])
            out[col] = torch.from_numpy(arr).float() if arr.dtype.kind == "f" else torch.from_numpy(arr)
        else:
            t = torch.from_numpy(vals)
            out[col] = t.float() if t.dtype == torch.float64 else t
    return out


class CSVDataset(DGLDataset):
    def __init__(self, data_path, force_reload=False, verbose=True, ndata_parser=None, edata_parser=None,
                 gdata_parser=None, transform=None):
        self.data_path = data_path
        meta_path = os.path.join(data_path, "meta.yaml")
        if not os.path.exists(meta_path):
            raise DGLError(f"'meta.yaml' cannot be found under {data_path}.")
        with open(meta_path) as f:
            self.meta = yaml.safe_load(f)
        self.graphs, self.data = [], {}
        super().__init__(self.meta.get("dataset_name", "csv_dataset"), raw_dir=os.path.dirname(data_path),
                         force_reload=force_reload, verbose=verbose, transform=transform)

    def process(self):
        sep = self.meta.get("separator", ",")
        node_frames, edge_frames = {}, {}
        for nd in self.meta.get("node_data", []):
            node_frames[nd.get("ntype", "_V")] = pd.read_csv(os.path.join(self.data_path, nd["file_name"]), sep=sep)
        for ed in self.meta.get("edge_data", []):
            et = tuple(ed.get("etype", ["_V", "_E", "_V"]))
            edge_frames[et] = pd.read_csv(os.path.join(self.data_path, ed["file_name"]), sep=sep)
        gids = set()
        for df in list(node_frames.values()) + list(edge_frames.values()):
            if "graph_id" in df.columns:
                gids.update(df["graph_id"].unique().tolist())
        gids = sorted(gids) if gids else [None]
        for gid in gids:
            data, nn, nfeat, efeat = {}, {}, {}, {}
            idmaps = {}
            for nt, df in node_frames.items():
                d = df if gid is None else df[df["graph_id"] == gid]
                ids = d["node_id"].to_numpy()
                idmaps[nt] = {v: i for i, v in enumerate(ids)}
                nn[nt] = len(ids)
                nfeat[nt] = _parse_feats(d, {"node_id", "graph_id"})
            for et, df in edge_frames.items():
                d = df if gid is None else df[df["graph_id"] == gid]
                s = [idmaps[et[0]][x] for x in d["src_id"].to_numpy()] if et[0] in idmaps else d["src_id"].to_numpy()
                t = [idmaps[et[2]][x] for x in d["dst_id"].to_numpy()] if et[2] in idmaps else d["dst_id"].to_numpy()
                data[et] = (torch.as_tensor(np.asarray(s, dtype=np.int64)), torch.as_tensor(np.asarray(t, dtype=np.int64)))
                efeat[et] = _parse_feats(d, {"src_id", "dst_id", "graph_id"})
            g = heterograph(data, nn or None)
            for nt, fs in nfeat.items():
                for k, v in fs.items():
                    g.nodes[nt].data[k] = v
            for et, fs in efeat.items():
                for k, v in fs.items():
                    g.edges[et].data[k] = v
            self.graphs.append(g)
        gd = self.meta.get("graph_data")
        if gd:
            df = pd.read_csv(os.path.join(self.data_path, gd["file_name"]), sep=sep)
            self.data = _parse_feats(df, {"graph_id"})

    def __getitem__(self, i):
        g = self.graphs[i] if self._transform is None else self._transform(self.graphs[i])
        if self.data:
            return g, {k: v[i] for k, v in self.data.items()}
        return g

    def __len__(self):
        return len(self.graphs)
