"""CSVDataset and its building blocks (reference tests/python/common/data/test_data.py::test_csvdataset*,
python/dgl/data/csv_dataset{,_base}.py)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch
import yaml

import dgl_b200 as dgl
from dgl_b200.data import CSVDataset
from dgl_b200.data import csv_dataset_base as B


def _write(d, meta, files):
    with open(os.path.join(d, "meta.yaml"), "w") as f:
        yaml.safe_dump(meta, f)
    for name, df in files.items():
        df.to_csv(os.path.join(d, name), index=False, sep=meta.get("separator", ","))


def test_single_homograph_with_feature_strings(tmp_path):
    d = str(tmp_path)
    nodes = pd.DataFrame({"node_id": [10, 30, 20], "label": [0, 1, 2], "feat": ["0.1, 0.2", "0.3, 0.4", "0.5, 0.6"]})
    edges = pd.DataFrame({"src_id": [10, 20, 30, 30], "dst_id": [20, 30, 10, 20], "w": [1.0, 2.0, 3.0, 4.0]})
    _write(d, {"dataset_name": "tiny", "node_data": [{"file_name": "nodes.csv"}], "edge_data": [{"file_name": "edges.csv"}]},
           {"nodes.csv": nodes, "edges.csv": edges})
    ds = CSVDataset(d)
    assert len(ds) == 1
    g = ds[0]
    # node index = row order in the CSV; edges refer to user ids
    assert g.num_nodes() == 3 and g.edges()[0].tolist() == [0, 2, 1, 1] and g.edges()[1].tolist() == [2, 1, 0, 2]
    assert g.ndata["feat"].dtype == torch.float32 and g.ndata["feat"].shape == (3, 2)
    assert torch.allclose(g.ndata["feat"][1], torch.tensor([0.3, 0.4]))
    assert g.ndata["label"].tolist() == [0, 1, 2] and g.edata["w"].tolist() == [1.0, 2.0, 3.0, 4.0]
    # cache written next to the CSVs and reloaded on the second construction
    assert os.path.exists(os.path.join(d, "tiny.bin"))
    os.remove(os.path.join(d, "edges.csv"))
    again = CSVDataset(d)
    assert again[0].num_edges() == 4
    with pytest.raises(Exception):
        CSVDataset(d, force_reload=True)


def test_multi_hetero_graphs_with_labels_and_custom_parsers(tmp_path):
    d = str(tmp_path)
    users = pd.DataFrame({"graph_id": [0, 0, 1, 1, 1], "node_id": ["a", "b", "a", "c", "d"], "age": [1, 2, 3, 4, 5]})
    items = pd.DataFrame({"graph_id": [0, 1], "node_id": [7, 9], "price": [0.5, 1.5]})
    buys = pd.DataFrame({"graph_id": [0, 0, 1], "src_id": ["b", "a", "d"], "dst_id": [7, 7, 9], "n": [1, 2, 3]})
    graphs = pd.DataFrame({"graph_id": [1, 0], "label": [11, 10], "score": [0.1, 0.9]})
    meta = {"dataset_name": "shop", "separator": "|",
            "node_data": [{"file_name": "u.csv", "ntype": "user"}, {"file_name": "i.csv", "ntype": "item"}],
            "edge_data": [{"file_name": "b.csv", "etype": ["user", "buys", "item"]}],
            "graph_data": {"file_name": "g.csv"}}
    _write(d, meta, {"u.csv": users, "i.csv": items, "b.csv": buys, "g.csv": graphs})

    calls = []

    def age_parser(df):
        calls.append(list(df.columns))
        return {"age2": df["age"].to_numpy() * 2}

    ds = CSVDataset(d, ndata_parser={"user": age_parser}, gdata_parser=lambda df: {"label": df["label"].to_numpy()})
    assert calls == [["age"]]          # id / graph-id columns are removed before the parser sees the frame
    assert len(ds) == 2
    g1, l1 = ds[0]                     # graph CSV order: graph id 1 first
    g0, l0 = ds[1]
    assert int(l1) == 11 and int(l0) == 10
    assert g0.num_nodes("user") == 2 and g1.num_nodes("user") == 3 and g0.num_nodes("item") == 1
    assert g0.nodes["user"].data["age2"].tolist() == [2, 4] and "age" not in g0.nodes["user"].data
    assert g0.nodes["item"].data["price"].tolist() == [0.5]
    u, v = g0.edges(etype="buys")
    assert u.tolist() == [1, 0] and v.tolist() == [0, 0]
    assert g1.edges(etype="buys")[0].tolist() == [2]
    sub = ds[torch.tensor([1])]
    assert len(sub) == 1 and int(sub[0][1]) == 10
    # several graph-level columns -> dict per item
    ds2 = CSVDataset(d, force_reload=True)
    g, lab = ds2[0]
    assert set(lab) == {"label", "score"} and abs(float(lab["score"]) - 0.1) < 1e-6


def test_meta_validation_errors(tmp_path):
    d = str(tmp_path)
    with pytest.raises(dgl.DGLError, match="meta.yaml"):
        CSVDataset(d)
    base = {"dataset_name": "x", "node_data": [{"file_name": "n.csv"}], "edge_data": [{"file_name": "e.csv"}]}
    for bad, pattern in [({**base, "version": "2.0.0"}, "version"),
                         ({**base, "oops": 1}, "unknown"),
                         ({k: v for k, v in base.items() if k != "dataset_name"}, "missing"),
                         ({**base, "node_data": [{"file_name": "a.csv"}, {"file_name": "b.csv"}]}, "duplicate node type"),
                         ({**base, "edge_data": [{"file_name": "a.csv"}, {"file_name": "b.csv"}]}, "duplicate edge type"),
                         ({**base, "edge_data": [{"file_name": "a.csv", "etype": ["a", "b"]}]}, "etype")]:
        with open(os.path.join(d, "meta.yaml"), "w") as f:
            yaml.safe_dump(bad, f)
        with pytest.raises(dgl.DGLError, match=pattern):
            B.load_yaml_with_sanity_check(os.path.join(d, "meta.yaml"))


def test_table_level_errors_and_constructor():
    with pytest.raises(dgl.DGLError, match="unique"):
        B.NodeData.to_dict([B.NodeData([1, 1], {})])
    with pytest.raises(dgl.DGLError, match="same length"):
        B.NodeData([1, 2], {"f": np.zeros(3)})
    nd = B.NodeData([5, 6, 7], {"h": np.arange(3.0)})
    with pytest.raises(dgl.DGLError, match="do not appear"):
        B.DGLGraphConstructor.construct_graphs(nd, B.EdgeData([5], [8], {}))
    graphs, data = B.DGLGraphConstructor.construct_graphs(nd, B.EdgeData([7, 5], [5, 6], {"w": np.array([1, 2])}))
    assert len(graphs) == 1 and data == {} and graphs[0].edges()[0].tolist() == [2, 0]
    assert graphs[0].ndata["h"].dtype == torch.float32
    with pytest.raises(dgl.DGLError, match="not in graph CSV"):
        B.DGLGraphConstructor.construct_graphs(nd, B.EdgeData([5], [6], {}), B.GraphData([3], {"y": np.array([1])}))
    # a graph id that only exists in the graph table yields an empty graph
    gs, data = B.DGLGraphConstructor.construct_graphs(nd, B.EdgeData([5], [6], {}), B.GraphData([0, 4], {"y": np.array([1, 2])}))
    assert gs[1].num_edges() == 0 and data["y"].shape == (2, 1)
    node_dict = B.NodeData.to_dict([nd])
    assert node_dict[0]["_V"]["mapping"][6] == 1 and len(node_dict[0]["_V"]["mapping"]) == 3
    with pytest.raises(KeyError):
        node_dict[0]["_V"]["mapping"][9]


def test_default_parser():
    df = pd.DataFrame({"a": [1, 2], "s": ["[1, 2]", "[3, 4]"], "Unnamed: 3": [0, 0]})
    with pytest.warns(Warning):
        out = B.DefaultDataParser()(df)
    assert set(out) == {"a", "s"} and out["s"].shape == (2, 2) and out["a"].tolist() == [1, 2]
