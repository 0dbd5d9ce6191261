"""GraphBolt internals: work splitting, datapipe-graph utilities, file helpers, PyG layer tuples
(reference python/dgl/graphbolt/{internal/*,datapipes/*,sampled_subgraph,feature_fetcher,base}.py; reference tests
tests/python/pytorch/graphbolt/internal/*, test_dataloader.py::test_find_dps/replace_dp)."""
import itertools
import json
import os

import numpy as np
import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.graphbolt as gb
from dgl_b200.graphbolt.datapipes import utils as U
from dgl_b200.graphbolt.datapipes import visualization as V
from dgl_b200.graphbolt.internal import (calculate_dir_hash, calculate_file_hash, calculate_range, check_dataset_change,
                                         compact_temporal_nodes, copy_or_convert_data, count_split, get_npy_dim,
                                         read_data, read_edges, save_data)


def test_count_split_hands_out_whole_batches():
    for total, workers, bs in itertools.product([0, 1, 7, 10, 33, 100], [1, 2, 3, 4], [1, 2, 4]):
        parts = [count_split(total, workers, w, bs) for w in range(workers)]
        assert sum(a for a, _ in parts) == total
        offs = [o for _, o in parts]
        assert offs == [sum(a for a, _ in parts[:w]) for w in range(workers)]          # prefix sums
        sizes = [a for a, _ in parts]
        assert max(sizes) - min(sizes) <= bs                                            # balanced to one batch
        # at most one worker holds a ragged (non multiple of bs) share
        assert sum(1 for a in sizes if a % bs) <= 1


def test_calculate_range_partitions_items_between_replicas_and_workers():
    for total, nr, nw, bs, dl, du in itertools.product([0, 5, 17, 64, 101], [1, 2, 3], [1, 2, 3], [1, 3, 4],
                                                       [False, True], [False, True]):
        spans = [[calculate_range(True, total, nr, r, nw, w, bs, dl, du) for w in range(nw)] for r in range(nr)]
        flat = [s for rep in spans for s in rep]
        # assigned ranges tile [0, total) without gaps or overlap
        pos = 0
        for start, assigned, emitted in flat:
            assert start == pos and 0 <= emitted <= assigned
            pos += assigned
        assert pos == total
        per_rep = [sum(e for _, _, e in rep) for rep in spans]
        if du:
            batches = [-(-e // bs) for e in per_rep]
            assert len(set(batches)) == 1                                              # aligned batch counts
        if dl:
            assert all(e % bs == 0 for e in per_rep)
    assert calculate_range(False, 10, 4, 1, 2, 0, 3, True, True) == (0, 10, 9)
    assert calculate_range(False, 10, 4, 1, 2, 0, 3, False, False) == (0, 10, 10)


def _pipeline():
    g = gb.from_dglgraph(dgl.rand_graph(100, 600))
    items = gb.ItemSet(torch.arange(50), names="seeds")
    sampler = gb.ItemSampler(items, batch_size=10)
    nb = sampler.sample_neighbor(g, [3, 3])
    return sampler, nb, nb.transform(lambda mb: mb)


def test_datapipe_graph_find_and_replace():
    sampler, nb, tail = _pipeline()
    graph = U.traverse_dps(tail)
    adj = U.datapipe_graph_to_adjlist(graph)
    assert len(adj) == 3 and adj[id(tail)][1] == [id(nb)] and adj[id(sampler)][1] == []
    assert U.find_dps(graph, gb.SubgraphSampler) == [nb]
    assert U.find_dps(graph, gb.ItemSampler) == [sampler]
    pre = U.PrefetcherIterDataPipe(sampler, 2)
    graph = U.replace_dp(graph, sampler, pre)
    assert nb.datapipe is pre and U.find_dps(graph, U.PrefetcherIterDataPipe) == [pre]
    batches = list(tail)
    assert len(batches) == 5 and torch.cat([b.seeds for b in batches]).tolist() == list(range(50))
    # replacing the sink returns the new sink's graph
    new_tail = tail.transform(lambda mb: mb)
    assert id(new_tail) in U.replace_dp(U.traverse_dps(tail), tail, new_tail)
    names = [n.name for n in V.to_nodes(tail)]
    assert names == ["ItemSampler", "Prefetcher", "NeighborSampler", "MiniBatchTransformer"]
    dot = str(V.to_graph(tail))
    assert '"ItemSampler" -> "Prefetcher"' in dot.replace("\t", "") or "ItemSampler -> Prefetcher" in dot


def test_prefetcher_propagates_errors_and_stops_early():
    class Boom:
        def __iter__(self):
            yield 1
            raise RuntimeError("boom")

        def __len__(self):
            return 2

    it = iter(U.PrefetcherIterDataPipe(Boom(), 1))
    assert next(it) == 1
    with pytest.raises(RuntimeError, match="boom"):
        next(it)
    with pytest.raises(ValueError):
        U.PrefetcherIterDataPipe([1], 0)
    slow = U.PrefetcherIterDataPipe(list(range(1000)), 2)
    for i, x in enumerate(slow):          # abandoning the iterator must not hang the producer thread
        if i == 3:
            break
    w = gb.MultiprocessingWrapper(gb.ItemSampler(gb.ItemSet(torch.arange(20), names="seeds"), batch_size=5), 2)
    assert len(list(w)) == 4
    assert len(list(gb.MultiprocessingWrapper(gb.ItemSampler(gb.ItemSet(torch.arange(20), names="seeds"), batch_size=5)))) == 4


def test_pyg_layer_tuples_homo_and_hetero():
    _, nb, _ = _pipeline()
    mb = next(iter(nb))
    x = torch.randn(mb.input_nodes.shape[0], 4)
    layer = mb.sampled_subgraphs[0].to_pyg(x)
    assert isinstance(layer, gb.PyGLayerData)
    n_dst = mb.sampled_subgraphs[0].sampled_csc.indptr.shape[0] - 1
    assert layer.size == (x.shape[0], n_dst) and layer.x[1].shape[0] == n_dst and layer.edge_index.shape[0] == 2
    assert int(layer.edge_index[1].max()) < n_dst and int(layer.edge_index[0].max()) < x.shape[0]
    sub = gb.SampledSubgraphImpl(sampled_csc={"u:r:v": gb.CSCFormatBase(indptr=torch.tensor([0, 1, 3]), indices=torch.tensor([2, 0, 1]))})
    feats = {"u": torch.randn(3, 2), "v": torch.randn(5, 2)}
    het = sub.to_pyg(feats)
    assert isinstance(het, gb.PyGLayerHeteroData) and het.size["u:r:v"] == (3, 2)
    assert het.edge_index["u:r:v"].tolist() == [[2, 0, 1], [0, 1, 1]] and het.x[1]["v"].shape[0] == 2


def test_small_helpers():
    assert gb.get_feature_key_list({"a": ["x", "y"]}, "node") == [("node", "a", "x"), ("node", "a", "y")]
    assert gb.get_feature_key_list(["f"], "edge") == [("edge", None, "f")] and gb.get_feature_key_list(None, "node") == []
    assert gb.is_scalar(3) and gb.is_scalar(torch.tensor(1)) and not gb.is_scalar(torch.ones(2))
    gb.load_graphbolt()
    _, nb, _ = _pipeline()
    assert gb.is_object_pinned(next(iter(nb))) is False
    n, t, c = compact_temporal_nodes([torch.tensor([5, 5]), torch.tensor([7])], [torch.tensor([1, 2]), torch.tensor([3])])
    assert n.tolist() == [5, 5, 7] and t.tolist() == [1, 2, 3] and [x.tolist() for x in c] == [[0, 1], [2]]
    nd, td, cd = compact_temporal_nodes({"a": [torch.tensor([1])]}, {"a": [torch.tensor([9])]})
    assert nd["a"].tolist() == [1] and cd["a"][0].tolist() == [0]
    s = gb.NeighborSamplerImpl(gb.ItemSampler(gb.ItemSet(torch.arange(4), names="seeds"), batch_size=2),
                               gb.from_dglgraph(dgl.rand_graph(10, 30)), [2])
    assert type(s).__name__ == "NeighborSampler"


def test_file_helpers(tmp_path):
    d = str(tmp_path)
    a = np.arange(12, dtype=np.int64)
    save_data(a, os.path.join(d, "raw", "a.npy"), "numpy")
    assert get_npy_dim(os.path.join(d, "raw", "a.npy")) == 1
    assert read_data(os.path.join(d, "raw", "a.npy"), "numpy").tolist() == a.tolist()
    assert read_data(os.path.join(d, "raw", "a.npy"), "numpy", in_memory=False).tolist() == a.tolist()
    save_data(torch.arange(6).view(2, 3), os.path.join(d, "raw", "t.pt"), "torch")
    assert read_data(os.path.join(d, "raw", "t.pt"), "torch").shape == (2, 3)
    with pytest.raises(RuntimeError):
        read_data("x", "csv")
    # 1-D feature becomes [N, 1]; ids are narrowed on request; unchanged arrays are copied verbatim
    copy_or_convert_data(os.path.join(d, "raw", "a.npy"), os.path.join(d, "out", "f.npy"), "numpy", is_feature=True)
    assert np.load(os.path.join(d, "out", "f.npy")).shape == (12, 1)
    copy_or_convert_data(os.path.join(d, "raw", "a.npy"), os.path.join(d, "out", "i.npy"), "numpy", within_int32=True)
    assert np.load(os.path.join(d, "out", "i.npy")).dtype == np.int32
    copy_or_convert_data(os.path.join(d, "raw", "a.npy"), os.path.join(d, "out", "c.npy"), "numpy")
    assert open(os.path.join(d, "out", "c.npy"), "rb").read() == open(os.path.join(d, "raw", "a.npy"), "rb").read()
    copy_or_convert_data(os.path.join(d, "raw", "t.pt"), os.path.join(d, "out", "t.npy"), "torch")
    assert np.load(os.path.join(d, "out", "t.npy")).shape == (2, 3)
    np.save(os.path.join(d, "e.npy"), np.array([[0, 1, 2], [1, 2, 0]]))
    src, dst = read_edges(d, "numpy", "e.npy")
    assert src.tolist() == [0, 1, 2] and dst.tolist() == [1, 2, 0]
    with open(os.path.join(d, "e.csv"), "w") as f:
        f.write("0,1\n2,3\n")
    src, dst = read_edges(d, "csv", "e.csv")
    assert src.tolist() == [0, 2] and dst.tolist() == [1, 3]
    import hashlib
    assert calculate_file_hash(os.path.join(d, "e.csv"), "sha1") == hashlib.sha1(b"0,1\n2,3\n").hexdigest()
    with pytest.raises(ValueError):
        calculate_file_hash(os.path.join(d, "e.csv"), "crc32")
    ds = os.path.join(d, "ds")
    os.makedirs(os.path.join(ds, "preprocessed"))
    with open(os.path.join(ds, "x.txt"), "w") as f:
        f.write("1")
    assert check_dataset_change(ds, "preprocessed")
    with open(os.path.join(ds, "preprocessed", "dataset_hash_value.txt"), "w") as f:
        json.dump(calculate_dir_hash(ds, ignore="dataset_hash_value.txt"), f)
    assert not check_dataset_change(ds, "preprocessed")
    with open(os.path.join(ds, "x.txt"), "w") as f:
        f.write("2")
    assert check_dataset_change(ds, "preprocessed")


# ----------------------------------------------------------------------------------------------- disk-backed features
@pytest.mark.parametrize("backend", ["uring", "pread"])
@pytest.mark.parametrize("dtype,width", [(np.float32, 37), (np.int64, 4), (np.float16, 1031)])
def test_disk_based_feature_backends(tmp_path, backend, dtype, width):
    import dgl_b200.graphbolt as gb
    from dgl_b200.base import DGLError
    rng = np.random.default_rng(0)
    table = (rng.random((5000, width)) * 100).astype(dtype)
    path = str(tmp_path / "feat.npy")
    np.save(path, table)
    try:
        feat = gb.DiskBasedFeature(path, io_backend=backend, queue_depth=16, num_threads=3)
    except DGLError as e:
        assert backend == "uring" and "io_uring" in str(e)
        pytest.skip("io_uring is not permitted in this environment")
    assert feat.io_backend == backend
    ids = torch.from_numpy(rng.integers(0, 5000, 3000))          # duplicates, unsorted, more rows than queue slots
    got = feat.read(ids)
    assert got.dtype == torch.from_numpy(table[:1]).dtype and tuple(got.shape) == (3000, width)
    assert np.array_equal(got.numpy(), table[ids.numpy()])
    assert feat.read(torch.tensor([], dtype=torch.int64)).shape == (0, width)
    assert np.array_equal(feat.read().numpy(), table)
    with pytest.raises(DGLError):
        feat.read(torch.tensor([5000]))
    assert feat.count() == 5000 and tuple(feat.size()) == (width,)


def test_disk_based_feature_auto_backend_env(tmp_path, monkeypatch):
    import dgl_b200.graphbolt as gb
    path = str(tmp_path / "f.npy")
    np.save(path, np.arange(40, dtype=np.float32).reshape(10, 4))
    monkeypatch.setenv("DGLB200_DISK_IO", "pread")
    assert gb.DiskBasedFeature(path).io_backend == "pread"
    monkeypatch.delenv("DGLB200_DISK_IO")
    assert gb.DiskBasedFeature(path).io_backend in ("uring", "pread")
