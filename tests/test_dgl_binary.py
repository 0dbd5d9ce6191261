"""The reference's binary graph / tensor containers (dgl_b200/data/dgl_binary.py): self round trips on every stored
format, index-table seeks, malformed input, and - when the reference build is installed under baseline/_ref - both
directions against the reference's own reader and writer (run in a subprocess; the two packages share module names)."""
import json
import os
import struct
import subprocess
import sys

import pytest
import torch

import dgl_b200 as dgl
from dgl_b200.base import DGLError
from dgl_b200.data import load_graphs, load_labels, save_graphs
from dgl_b200.data import dgl_binary
from dgl_b200.data.compat import load_tensors, save_tensors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _graphs():
    g = dgl.graph(([0, 1, 2, 2, 3], [1, 2, 3, 0, 3]), num_nodes=6)
    g.ndata["x"] = torch.randn(6, 3)
    g.ndata["b"] = torch.randn(6, 2).bfloat16()
    g.edata["w"] = torch.arange(5, dtype=torch.float64)
    hg = dgl.heterograph({("u", "f", "i"): ([0, 1], [1, 2]), ("i", "r", "u"): ([2, 0, 1], [0, 0, 1]),
                          ("u", "k", "u"): ([0], [1])}, idtype=torch.int32)
    hg.nodes["u"].data["h"] = torch.randn(2, 2)
    hg.edges["r"].data["m"] = torch.tensor([3, 1, 2], dtype=torch.int16)
    return g, hg


def _same(a, b):
    assert a.ntypes == b.ntypes and a.canonical_etypes == b.canonical_etypes and a.idtype == b.idtype
    for nt in a.ntypes:
        assert a.num_nodes(nt) == b.num_nodes(nt)
        assert set(a.nodes[nt].data.keys()) == set(b.nodes[nt].data.keys())
        for k in a.nodes[nt].data:
            assert torch.equal(a.nodes[nt].data[k], b.nodes[nt].data[k])
    for c in a.canonical_etypes:
        for x, y in zip(a.edges(etype=c), b.edges(etype=c)):
            assert torch.equal(x, y)
        for k in a.edges[c].data:
            assert torch.equal(a.edges[c].data[k], b.edges[c].data[k])


@pytest.mark.parametrize("formats", [None, "coo", "csr", "csc", ["csr", "csc"], ["coo", "csr", "csc"]])
def test_round_trip_every_format(tmp_path, formats):
    g, hg = _graphs()
    path = str(tmp_path / "graphs.bin")
    save_graphs(path, [g, hg], {"y": torch.tensor([4, 2])}, formats=formats)
    assert dgl_binary.is_dgl_binary(path)
    out, labels = load_graphs(path)
    assert labels["y"].tolist() == [4, 2] and load_labels(path)["y"].tolist() == [4, 2]
    _same(g, out[0])
    _same(hg, out[1])
    only, _ = load_graphs(path, [1])
    _same(hg, only[0])
    both, _ = load_graphs(path, [1, 0])
    _same(g, both[1])


def test_loaded_tensors_are_writable_and_file_is_untouched(tmp_path):
    g, _ = _graphs()
    path = str(tmp_path / "g.bin")
    save_graphs(path, [g])
    before = open(path, "rb").read()
    out, _ = load_graphs(path)
    out[0].ndata["x"] += 1.0  # copy-on-write mapping
    assert open(path, "rb").read() == before
    assert torch.allclose(out[0].ndata["x"], g.ndata["x"] + 1.0)


def test_tensor_container_round_trip(tmp_path):
    path = str(tmp_path / "t.dgl")
    d = {"a": torch.randn(3, 4), "idx": torch.arange(7, dtype=torch.int32), "h": torch.randn(2, 2, 2).half(),
         "empty": torch.empty(0, 5)}
    save_tensors(path, d)
    back = load_tensors(path)
    assert set(back) == set(d)
    for k in d:
        assert back[k].dtype == d[k].dtype and back[k].shape == d[k].shape and torch.equal(back[k], d[k])


def test_blocks_and_batches_keep_the_archive_container(tmp_path):
    g, _ = _graphs()
    bg = dgl.batch([g, g])
    path = str(tmp_path / "b.bin")
    save_graphs(path, [bg])
    assert not dgl_binary.is_dgl_binary(path)
    out, _ = load_graphs(path)
    assert out[0].batch_size == 2
    with pytest.raises(DGLError):
        save_graphs(path, [bg], container="dgl")


def test_malformed_files_raise_dglerror(tmp_path):
    g, _ = _graphs()
    path = str(tmp_path / "g.bin")
    save_graphs(path, [g])
    blob = open(path, "rb").read()
    cut = str(tmp_path / "cut.bin")
    open(cut, "wb").write(blob[: 4096 + 40])
    with pytest.raises(DGLError):
        load_graphs(cut)
    bad = bytearray(blob)
    struct.pack_into("<Q", bad, 8, 7)  # version field
    badp = str(tmp_path / "ver.bin")
    open(badp, "wb").write(bytes(bad))
    with pytest.raises(DGLError):
        load_graphs(badp)
    with pytest.raises(DGLError):
        load_graphs(path, [5])


_REF_SCRIPT = r"""
import json, sys, torch
sys.path.insert(0, sys.argv[1])
import dgl
mode, path = sys.argv[2], sys.argv[3]
if mode == "write":
    g = dgl.graph(([0, 1, 2, 2, 3], [1, 2, 3, 0, 3]), num_nodes=6)
    g.ndata["x"] = torch.arange(18, dtype=torch.float32).reshape(6, 3)
    g.edata["w"] = torch.arange(5, dtype=torch.float64)
    hg = dgl.heterograph({("u", "f", "i"): ([0, 1], [1, 2]), ("i", "r", "u"): ([2, 0, 1], [0, 0, 1]),
                          ("u", "k", "u"): ([0], [1])}, idtype=torch.int32)
    hg.nodes["u"].data["h"] = torch.ones(2, 2)
    dgl.save_graphs(path, [g, hg], {"y": torch.tensor([4, 2])}, formats=sys.argv[4].split(",") if sys.argv[4] else None)
    dgl.data.utils.save_tensors(path + ".t", {"a": torch.arange(6.).reshape(2, 3), "i": torch.arange(4)})
else:
    gs, labels = dgl.load_graphs(path)
    out = {"labels": {k: v.tolist() for k, v in labels.items()}, "graphs": []}
    for g in gs:
        out["graphs"].append({
            "ntypes": g.ntypes, "cet": [list(c) for c in g.canonical_etypes], "idtype": str(g.idtype),
            "num_nodes": [g.num_nodes(t) for t in g.ntypes],
            "edges": [[t.tolist() for t in g.edges(etype=c)] for c in g.canonical_etypes],
            "ndata": {t: {k: v.float().tolist() for k, v in g.nodes[t].data.items()} for t in g.ntypes},
            "edata": {c[1]: {k: v.float().tolist() for k, v in g.edges[c].data.items()} for c in g.canonical_etypes}})
    t = dgl.data.utils.load_tensors(path + ".t")
    out["tensors"] = {k: v.tolist() for k, v in t.items()}
    print("RESULT " + json.dumps(out))
"""


def _ref(*args):
    env = dict(os.environ, DGLBACKEND="pytorch", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", _REF_SCRIPT, REF, *args], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "dgl", "libdgl.so")),
                               reason="reference build not installed (script/build_reference.sh)")


@needs_ref
@pytest.mark.parametrize("formats", ["", "csc", "coo,csr"])
def test_files_written_by_the_reference_load_here(tmp_path, formats):
    path = str(tmp_path / "ref.bin")
    _ref("write", path, formats)
    gs, labels = load_graphs(path)
    assert labels["y"].tolist() == [4, 2]
    g, hg = gs
    assert g.num_nodes() == 6 and [t.tolist() for t in g.edges()] == [[0, 1, 2, 2, 3], [1, 2, 3, 0, 3]]
    assert torch.equal(g.ndata["x"], torch.arange(18, dtype=torch.float32).reshape(6, 3))
    assert torch.equal(g.edata["w"], torch.arange(5, dtype=torch.float64))
    assert hg.idtype == torch.int32 and hg.ntypes == ["i", "u"]
    assert sorted(hg.canonical_etypes) == [("i", "r", "u"), ("u", "f", "i"), ("u", "k", "u")]
    assert [t.tolist() for t in hg.edges(etype="r")] == [[2, 0, 1], [0, 0, 1]]
    assert torch.equal(hg.nodes["u"].data["h"], torch.ones(2, 2))
    t = load_tensors(path + ".t")
    assert t["a"].tolist() == [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]] and t["i"].tolist() == [0, 1, 2, 3]


@needs_ref
@pytest.mark.parametrize("formats", [None, "csr", ["coo", "csc"]])
def test_files_written_here_load_in_the_reference(tmp_path, formats):
    g, hg = _graphs()
    del g.ndata["b"]  # the reference's torch bridge has no bfloat16 NDArray -> tensor conversion on CPU builds
    path = str(tmp_path / "ours.bin")
    save_graphs(path, [g, hg], {"y": torch.tensor([4, 2])}, formats=formats)
    save_tensors(path + ".t", {"a": torch.arange(6.).reshape(2, 3), "i": torch.arange(4)})
    out = _ref("read", path)
    rec = json.loads([l for l in out.splitlines() if l.startswith("RESULT ")][0][7:])
    assert rec["labels"]["y"] == [4, 2]
    assert rec["tensors"] == {"a": [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]], "i": [0, 1, 2, 3]}
    for ours, theirs in zip((g, hg), rec["graphs"]):
        assert theirs["ntypes"] == ours.ntypes
        assert [tuple(c) for c in theirs["cet"]] == ours.canonical_etypes
        assert theirs["num_nodes"] == [ours.num_nodes(t) for t in ours.ntypes]
        assert theirs["idtype"] == str(ours.idtype)
        for c, e in zip(ours.canonical_etypes, theirs["edges"]):
            assert e == [t.tolist() for t in ours.edges(etype=c)]
        for t in ours.ntypes:
            for k, v in ours.nodes[t].data.items():
                assert theirs["ndata"][t][k] == v.float().tolist()
        for c in ours.canonical_etypes:
            for k, v in ours.edges[c].data.items():
                assert theirs["edata"][c[1]][k] == v.float().tolist()


_REF_PARTITION = r"""
import sys, torch, unittest.mock as um
sys.path.insert(0, sys.argv[1])
import dgl
# GraphBolt's C++ library is not part of the reference build used here; dgl.distributed only needs the name to import
_gb = um.MagicMock(); _gb.FusedCSCSamplingGraph = type("FusedCSCSamplingGraph", (), {}); sys.modules["dgl.graphbolt"] = _gb
import dgl.distributed
torch.manual_seed(0)
g = dgl.rand_graph(200, 1500)
g.ndata["feat"] = torch.arange(200, dtype=torch.float32).unsqueeze(1).repeat(1, 4)
g.edata["w"] = torch.arange(1500, dtype=torch.float32)
dgl.distributed.partition_graph(g, "tg", 2, sys.argv[2], part_method="random", num_hops=1)
torch.save({"src": g.edges()[0], "dst": g.edges()[1]}, sys.argv[2] + "/orig.pt")
"""


@needs_ref
def test_partitions_written_by_the_reference_serve_a_distgraph(tmp_path):
    import dgl_b200.distributed as dist
    out = str(tmp_path / "parts")
    env = dict(os.environ, DGLBACKEND="pytorch", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", _REF_PARTITION, REF, out], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    cfg = os.path.join(out, "tg.json")
    total_inner = 0
    for p in (0, 1):
        g, nf, ef, gpb, name, ntypes, etypes = dist.load_partition(cfg, p)
        assert name == "tg" and ntypes == ["_N"] and {"_ID", "inner_node", "part_id"} <= set(g.ndata.keys())
        inner = g.ndata["inner_node"].bool()
        total_inner += int(inner.sum())
        assert nf["_N/feat"].shape == (int(inner.sum()), 4) and ef["_N:_E:_N/w"].shape[0] == int(g.edata["inner_edge"].sum())
    assert total_inner == 200
    dist.initialize(None)
    dg = dist.DistGraph("tg", part_config=cfg)
    assert dg.num_nodes() == 200 and dg.num_edges() == 1500
    rows = dg.ndata["feat"][torch.arange(200)]
    # the reference relabels nodes; every original row appears exactly once
    assert sorted(rows[:, 0].tolist()) == list(range(200))
    fr = dist.sample_neighbors(dg, torch.arange(50), 3)
    assert 0 < fr.num_edges() <= 150
