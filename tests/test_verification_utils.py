"""tools/verification_utils.py: every check passes on real partitions (homogeneous and heterogeneous) and fires on a
corrupted one."""
import json
import os
import sys

import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.distributed as dd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import verification_utils as vu  # noqa: E402
from files import setdir  # noqa: E402


def _hetero():
    torch.manual_seed(0)
    g = dgl.heterograph({("u", "f", "i"): (torch.randint(0, 60, (400,)), torch.randint(0, 90, (400,))),
                         ("i", "r", "u"): (torch.randint(0, 90, (300,)), torch.randint(0, 60, (300,)))},
                        num_nodes_dict={"u": 60, "i": 90})
    g.nodes["u"].data["h"] = torch.randn(60, 4)
    g.nodes["i"].data["x"] = torch.arange(90, dtype=torch.float32).unsqueeze(1)
    g.edges["f"].data["w"] = torch.arange(400, dtype=torch.float32)
    return g


def test_checks_pass_on_hetero_partitions_and_catch_corruption(tmp_path):
    g = _hetero()
    nmap, emap = dd.partition_graph(g, "hg", 3, str(tmp_path), num_hops=1, return_mapping=True)
    cfg = str(tmp_path / "hg.json")
    schema = vu.load_part_schema(cfg)
    for p in range(3):
        pg, nf, ef, gpb, name, ntypes, etypes = dd.load_partition(cfg, p)
        vu.verify_partition_data_types(pg, p)
        vu.verify_partition_formats(pg, "coo", p)
        vu.verify_metadata_counts(schema, pg, p)
        vu.verify_graph_feats(g, gpb, pg, nf, ef, orig_nids=nmap, orig_eids={":".join(c): v for c, v in emap.items()}, partid=p)
    # corrupt one stored feature row: the feature check must notice
    pg, nf, ef, gpb, *_ = dd.load_partition(cfg, 1)
    key = next(k for k in nf if k.endswith("/x") and nf[k].shape[0])
    nf[key] = nf[key].clone()
    nf[key][0] += 1.0
    with pytest.raises(AssertionError, match="node feature"):
        vu.verify_graph_feats(g, gpb, pg, nf, ef, orig_nids=nmap, orig_eids={":".join(c): v for c, v in emap.items()}, partid=1)
    bad = json.loads(json.dumps(schema))
    bad["node_map"]["u"][0][1] += 1
    with pytest.raises(AssertionError, match="inner nodes"):
        vu.verify_metadata_counts(bad, dd.load_partition(cfg, 0)[0], 0)
    with pytest.raises(AssertionError, match="sparse format"):
        g2 = dgl.graph(([0, 1], [1, 0])).formats("csr")
        vu.verify_partition_formats(g2, "coo", 0)
    pg.ndata["inner_node"] = pg.ndata["inner_node"].float()
    with pytest.raises(AssertionError, match="dtype"):
        vu.verify_partition_data_types(pg, 1)


def test_node_partition_ids_and_helpers(tmp_path):
    g = dgl.rand_graph(120, 900)
    order, _ = dd.partition_graph(g, "g", 2, str(tmp_path), num_hops=1, return_mapping=True)
    cfg = str(tmp_path / "g.json")
    owner = torch.empty(120, dtype=torch.int64)
    gpbs = []
    for p in range(2):
        pg, _, _, gpb, *_ = dd.load_partition(cfg, p)
        inner = pg.ndata["_ID"][pg.ndata["inner_node"].bool()]
        owner[order[inner]] = p
        gpbs.append((pg, gpb))
    with setdir(str(tmp_path / "assign")):
        with open("_N.txt", "w") as f:
            f.write("\n".join(str(int(x)) for x in owner))
    assert os.getcwd() != str(tmp_path / "assign")
    ids = vu.get_node_partids(str(tmp_path / "assign"), {"node_type": ["_N"], "num_nodes_per_type": [120]})
    for p, (pg, gpb) in enumerate(gpbs):
        vu.verify_node_partitionids(ids, pg, g, gpb, orig_nids={"_N": order}, partid=p)
    ids["_N"] = 1 - ids["_N"]
    with pytest.raises(AssertionError, match="assigned elsewhere"):
        vu.verify_node_partitionids(ids, gpbs[0][0], g, gpbs[0][1], orig_nids={"_N": order}, partid=0)
    import numpy as np
    np.save(str(tmp_path / "a.npy"), np.arange(4))
    assert vu.read_file(str(tmp_path / "a.npy")).tolist() == [0, 1, 2, 3]
    np.savetxt(str(tmp_path / "a.csv"), np.arange(6).reshape(3, 2), delimiter=",")
    assert vu.read_file(str(tmp_path / "a.csv")).shape == (3, 2)
    with pytest.raises(ValueError):
        vu.read_file(str(tmp_path / "a.xyz"))
