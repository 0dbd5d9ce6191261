"""Small reference-path modules: mpops.edgewise, multiprocessing.thread_wrapped_func, frame.Scheme, function.TargetCode,
heterograph helpers, udf/ndarray/container, sampling.EidExcluder, ops.reshape_lhs_rhs, storages registry."""
import os
import pickle

import numpy as np
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200 import mpops, storages
from dgl_b200.frame import Frame, Scheme, infer_scheme
from dgl_b200.heterograph import combine_frames, find_src_dst_ntypes, make_canonical_etypes, pad_tuple
from dgl_b200.multiprocessing import thread_wrapped_func
from dgl_b200.sampling import EidExcluder


def test_mpops_edgewise_with_etype():
    g = dgl.heterograph({("u", "f", "u"): ([0, 1, 2], [2, 3, 4]), ("u", "l", "m"): ([3, 3, 1, 2], [0, 0, 1, 1])})
    x, y = torch.randn(5, 5), torch.randn(2, 5)
    assert torch.equal(mpops.copy_u(g, x, etype="l"), x[[3, 3, 1, 2]])
    assert torch.equal(mpops.copy_v(g, y, etype="l"), y[[0, 0, 1, 1]])
    assert torch.allclose(mpops.u_add_v(g, x, y, etype="l"), x[[3, 3, 1, 2]] + y[[0, 0, 1, 1]])
    assert torch.allclose(mpops.u_mul_v(g, x, x, etype="f"), x[[0, 1, 2]] * x[[2, 3, 4]])
    h = dgl.rand_graph(10, 20)
    a, b = torch.randn(10, 3), torch.randn(10, 3)
    u, v = h.edges()
    assert torch.allclose(mpops.u_dot_v(h, a, b).view(-1), (a[u] * b[v]).sum(-1), atol=1e-6)
    assert torch.allclose(mpops.u_sub_v(h, a, b), a[u] - b[v]) and torch.allclose(mpops.u_div_v(h, a, b), a[u] / b[v])
    assert set(mpops.edgewise.__all__) == {"copy_u", "copy_v", "u_add_v", "u_sub_v", "u_mul_v", "u_div_v", "u_dot_v"}


def test_thread_wrapped_func():
    assert thread_wrapped_func(lambda a: a * 2)(4) == 8
    with pytest.raises(ZeroDivisionError, match="Traceback"):
        thread_wrapped_func(lambda: 1 / 0)()


def test_scheme_and_target_code():
    s = infer_scheme(torch.zeros(3, 4))
    assert s == ((4,), torch.float32) and s.shape == (4,) and s.dtype == torch.float32
    assert pickle.loads(pickle.dumps(s)) == s and isinstance(pickle.loads(pickle.dumps(s)), Scheme)
    assert Frame({"a": torch.zeros(2, 3)}).schemes["a"].shape == (3,)
    T = dgl.function.TargetCode
    assert (T.SRC, T.DST, T.EDGE) == (0, 1, 2) and T.CODE2STR[2] == "e"


def test_heterograph_helpers():
    g = dgl.heterograph({("u", "f", "u"): ([0], [1]), ("u", "l", "m"): ([0], [0])})
    mg = g._graph.metagraph
    assert make_canonical_etypes(g.etypes, g.ntypes, mg) == g.canonical_etypes
    with pytest.raises(dgl.DGLError):
        make_canonical_etypes(["f"], g.ntypes, mg)
    assert find_src_dst_ntypes(g.ntypes, mg) is None
    b = dgl.heterograph({("a", "r", "b"): ([0], [0])})
    assert find_src_dst_ntypes(b.ntypes, b._graph.metagraph) == ({"a": 0}, {"b": 1})
    assert pad_tuple(1, 3) == (1, None, None) and pad_tuple((1, 2), 2) == (1, 2) and pad_tuple((1, 2, 3), 2) is None
    frames = [Frame({"a": torch.zeros(2), "b": torch.ones(2)}), Frame({"a": torch.ones(3)}), Frame({}, num_rows=0)]
    out = combine_frames(frames, [0, 1, 2])
    assert list(out.keys()) == ["a"] and out["a"].tolist() == [0, 0, 1, 1, 1]
    assert combine_frames([Frame({"a": torch.zeros(2)}), Frame({"b": torch.zeros(2)})], [0, 1]) is None
    with pytest.raises(dgl.DGLError):
        combine_frames([Frame({"a": torch.zeros(2)}), Frame({"a": torch.zeros(2, 3)})], [0, 1])
    assert list(combine_frames(frames, [0, 1], col_names=["a"]).keys()) == ["a"]


def test_reference_path_modules():
    assert dgl.udf.EdgeBatch is dgl.core.EdgeBatch and dgl.udf.NodeBatch is dgl.core.NodeBatch
    nd = dgl.ndarray
    assert nd.array([1, 2]).tolist() == [1, 2] and nd.cpu() == torch.device("cpu") and nd.gpu(1) == torch.device("cuda", 1)
    arr = np.arange(3)
    t = nd.zerocopy_from_numpy(arr)
    arr[0] = 7
    assert t[0] == 7 and nd.NDArray is torch.Tensor
    assert nd.cast_to_signed(torch.tensor([255], dtype=torch.uint8)).tolist() == [-1]
    assert nd.from_dlpack(nd.to_dlpack(torch.arange(3))).tolist() == [0, 1, 2]
    assert nd.SparseFormat.FORMAT2STR[2] == "CSR" and nd.SparseMatrix(2, [1, 2], [1]).flags == [True]
    name = "dglb200_nd_test"
    a = nd.create_shared_mem_array(name, (2,), torch.int64)
    assert nd.exist_shared_mem_array(name) and nd.get_shared_mem_array(name, (2,), torch.int64).shape == (2,)
    from dgl_b200.utils.shared_mem import unlink
    del a
    unlink(name)
    c = dgl.container
    assert c.List([1, 2])[1] == 2 and len(c.List([1, 2, 3])[0:2]) == 2 and dict(c.Map({"a": 1})) == {"a": 1}
    with pytest.raises(IndexError):
        c.List([1])[3]
    with pytest.raises(TypeError):
        c.StrMap({1: 2})
    assert c.convert_to_strmap({"k": 1})["k"] == 1 and c.Value(3).data == 3


def test_eid_excluder_homo_and_hetero():
    g = dgl.rand_graph(20, 100)
    fr = dgl.sampling.sample_neighbors(g, torch.arange(20), 3)
    banned = fr.edata[dgl.EID][:5].clone()
    w = [torch.rand(fr.num_edges())]
    w0 = w[0].clone()
    out = EidExcluder(banned)(fr, w)
    assert out.num_edges() == fr.num_edges() - 5 and not torch.isin(out.edata[dgl.EID], banned).any()
    assert w[0].shape[0] == out.num_edges() and torch.equal(w[0], w0[5:])
    # kept edges still point at the right parent edges
    pu, pv = g.find_edges(out.edata[dgl.EID])
    ou, ov = out.edges()
    assert torch.equal(pu, ou) and torch.equal(pv, ov)
    same = EidExcluder(torch.tensor([10 ** 6]))(fr)
    assert same.num_edges() == fr.num_edges()
    hg = dgl.heterograph({("a", "r", "b"): ([0, 1, 2], [0, 0, 1]), ("b", "s", "a"): ([0, 1], [1, 2])})
    hfr = dgl.sampling.sample_neighbors(hg, {"a": [1, 2], "b": [0, 1]}, -1)
    out = EidExcluder({"r": torch.tensor([1])})(hfr)
    assert out.num_edges("r") == 2 and out.num_edges("s") == 2 and 1 not in out.edges["r"].data[dgl.EID].tolist()


def test_reshape_lhs_rhs_and_storage_registry():
    a, b = dgl.ops.reshape_lhs_rhs(torch.zeros(4, 3, 1), torch.zeros(6, 5, 3, 4))
    assert a.shape == (4, 1, 3, 1) and b.shape == (6, 5, 3, 4)
    a, b = dgl.ops.reshape_lhs_rhs(torch.zeros(4, 3), torch.zeros(6, 3))
    assert a.shape == (4, 3)
    t = storages.wrap_storage(torch.arange(10.0).view(5, 2))
    assert isinstance(t, storages.TensorStorage) and t.fetch(torch.tensor([3, 0]), "cpu").tolist() == [[6.0, 7.0], [0.0, 1.0]]
    n = storages.wrap_storage(np.arange(6).reshape(3, 2))
    assert n.fetch_async(torch.tensor([2]), "cpu").wait().tolist() == [[4, 5]]
    assert storages.ThreadedFuture(lambda x, y: x + y, (1, 2)).wait() == 3
    with pytest.raises(ZeroDivisionError):
        storages.ThreadedFuture(lambda: 1 / 0, ()).wait()

    @storages.register_storage_wrapper(list)
    class ListStorage(storages.FeatureStorage):
        def __init__(self, x):
            self.x = x

        def fetch(self, indices, device, pin_memory=False, **kwargs):
            return [self.x[i] for i in indices]

    try:
        assert storages.wrap_storage([1, 2, 3]).fetch([0, 2], "cpu") == [1, 3]
    finally:
        del storages.STORAGE_WRAPPERS[list]
    with pytest.raises(AssertionError):
        storages.wrap_storage("nope")
    assert storages.BaseTensorStorage(torch.arange(4)).fetch(torch.tensor([1]), "cpu").tolist() == [1]


def test_method_level_additions():
    from dgl_b200 import graphbolt as gb
    from dgl_b200.transforms import module as TM
    from dgl_b200.graphbolt.datapipes.visualization import Node
    m = torch.rand(20, 20)
    eps = TM.GDC([0.5, 0.3], avg_degree=5).get_eps(20, m)
    assert int((m >= eps).sum()) == 100 and TM.PPR(eps=0.1).get_eps(20, m) == 0.1
    assert TM.HeatKernel(avg_degree=50).get_eps(20, m) == float("-inf")
    s = dgl.dataloading.LaborSampler([2, 2])
    g = dgl.rand_graph(100, 1000)
    s.set_seed(7)
    a = s.sample(g, torch.arange(10))[2][0].edata[dgl.EID]
    s.set_seed(7)
    b = s.sample(g, torch.arange(10))[2][0].edata[dgl.EID]
    assert torch.equal(a, b) and s.cnt == 1
    csc = gb.CSCFormatBase(indptr=torch.tensor([0, 1]), indices=torch.tensor([0]))
    assert csc.is_pinned() is False
    cache = gb.CPUFeatureCache((4, 2), torch.float32)
    table = torch.arange(20.0).view(10, 2)
    reads = []

    def reader(keys):
        reads.append(keys.tolist())
        return table[keys]

    assert torch.equal(cache.query_and_replace(torch.tensor([1, 3]), reader), table[[1, 3]])
    assert torch.equal(cache.query_and_replace(torch.tensor([3, 5]), reader), table[[3, 5]])
    assert reads == [[1, 3], [5]] and cache.max_size_in_bytes == 4 * 8 and cache.is_pinned() is False
    a_, b_ = Node(object(), "a"), Node(object(), "b")
    a_.add_child(b_)
    assert a_ in b_.parents and b_ in a_.childs
    a_.remove_child(b_)
    assert not b_.parents and not a_.childs
    conv = dgl.nn.RelGraphConv(4, 3, 2)
    before = conv.loop_weight.clone()
    conv.reset_parameters()
    assert not torch.equal(before, conv.loop_weight)
    mm = dgl.nn.GMMConv(3, 2, 2, 2)
    mm.set_allow_zero_in_degree(True)
    assert mm._allow_zero_in_degree is True
    dgl.nn.GatedGraphConv(3, 3, 1, 1).set_allow_zero_in_degree(True)
    part = dgl.partition.NDArrayPartition(10, 2)
    assert part.get() is part
    emb = dgl.nn.NodeEmbedding(6, 2, "mm_emb")
    opt = dgl.optim.SparseAdagrad([emb], lr=1.0)
    opt.setup([emb])
    w0 = emb.weight.clone()
    opt.update(torch.tensor([1, 1, 4]), torch.ones(3, 2), emb)
    assert not torch.equal(emb.weight[1], w0[1]) and torch.equal(emb.weight[0], w0[0])
    from dgl_b200.distributed import kvstore
    pol = kvstore.PartitionPolicy("edge:a:r:b", None)
    assert pol.type_name == ("a", "r", "b") and not pol.is_node() and str(pol.get_data_name("w")) == "edge~a:r:b~w"
    assert kvstore.PartitionPolicy("node:user", None).type_name == "user"


def test_graph_archive_is_tensor_only_and_rejects_foreign_files(tmp_path):
    """loading a graph file must not unpickle arbitrary objects; malformed input raises DGLError."""
    import pickle
    import dgl_b200 as dgl
    g = dgl.heterograph({("a", "r", "b"): ([0, 1, 2], [1, 0, 1]), ("b", "s", "a"): ([0, 1], [2, 2])})
    g.nodes["a"].data["x"] = torch.randn(3, 4)
    p = str(tmp_path / "g.pt")
    dgl.save_graphs(p, [g, dgl.batch([dgl.rand_graph(4, 6), dgl.rand_graph(3, 5)])], {"y": torch.tensor([1, 2])})
    torch.load(p, weights_only=True)                                   # the archive holds tensors + plain containers only
    gs, labels = dgl.load_graphs(p)
    assert gs[0].canonical_etypes == g.canonical_etypes and torch.equal(gs[0].nodes["a"].data["x"], g.nodes["a"].data["x"])
    assert gs[1].batch_size == 2 and labels["y"].tolist() == [1, 2]
    with pytest.raises(dgl.DGLError):
        dgl.load_graphs(p, [5])

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned > %s" % (tmp_path / "pwned"),))

    bad = str(tmp_path / "evil.pt")
    with open(bad, "wb") as fh:
        pickle.dump({"magic": "DGLB200-GRAPHS", "graphs": [Evil()]}, fh)
    with pytest.raises(dgl.DGLError):
        dgl.load_graphs(bad)
    assert not (tmp_path / "pwned").exists()
    junk = str(tmp_path / "junk.bin")
    with open(junk, "wb") as fh:
        fh.write(b"\x3f\xa1\xb4\x46\xf0\x4f\x2e\xdd" + b"\0" * 64)       # libdgl magic, truncated body: parsed natively
    with pytest.raises(dgl.DGLError, match="version|truncated|too short"):
        dgl.load_graphs(junk)


def test_dataloader_worker_processes_and_early_exit():
    """num_workers > 0 on a host graph = real sampling processes (ordered results, prefetched features); abandoning an
    iterator - process pool or prefetch thread - neither hangs nor leaks the producer."""
    import threading
    import dgl_b200 as dgl
    from dgl_b200.dataloading import DataLoader, NeighborSampler
    g = dgl.rand_graph(3000, 30000)
    g.ndata["feat"] = torch.randn(3000, 8)
    g.ndata["label"] = torch.randint(0, 4, (3000,))
    s = NeighborSampler([4, 4], prefetch_node_feats=["feat"], prefetch_labels=["label"])
    dl = DataLoader(g, torch.arange(1000), s, batch_size=100, num_workers=2)
    assert dl._use_processes
    seen = []
    for inp, out, blocks in dl:
        seen.append(out)
        assert torch.equal(blocks[0].srcdata["feat"], g.ndata["feat"][inp])
        assert torch.equal(blocks[-1].dstdata["label"], g.ndata["label"][out])
    assert torch.equal(torch.cat(seen), torch.arange(1000))              # same order as the single-process loader
    assert dl._pool is None                                                # pool shut down at the end of the epoch
    it = iter(DataLoader(g, torch.arange(1000), s, batch_size=100, num_workers=2))
    next(it)
    it.close()                                                             # early exit: workers are stopped
    before = threading.active_count()
    os.environ["DGLB200_LOADER_THREADS"] = "1"
    try:
        dl2 = DataLoader(g, torch.arange(1000), s, batch_size=100, num_workers=1)
        assert not dl2._use_processes and dl2.use_prefetch_thread
        it2 = iter(dl2)
        next(it2)
        it2.close()                                                        # the producer thread must notice and exit
    finally:
        del os.environ["DGLB200_LOADER_THREADS"]
    import time
    for _ in range(50):
        if threading.active_count() <= before:
            break
        time.sleep(0.1)
    assert threading.active_count() <= before


def _forked_native_call(q):
    import torch
    import dgl_b200 as dgl
    g = dgl.rand_graph(20000, 200000)
    q.put(float(dgl.ops.copy_u_sum(g, torch.ones(20000, 8)).sum()))


def test_native_kernels_survive_fork_after_parallel_regions():
    """A child forked after the parent has run multi-threaded native kernels must not deadlock in OpenMP: the loader
    pins forked children to one-thread teams (GNU OpenMP keeps no threads across fork)."""
    import multiprocessing as mp
    import dgl_b200 as dgl
    big = dgl.rand_graph(100000, 1000000)
    x = torch.randn(100000, 16)
    for _ in range(3):                                        # make sure the parent's OpenMP pool exists
        dgl.ops.copy_u_sum(big, x)
    dgl.sampling.sample_neighbors(big, torch.arange(20000), 5)
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_forked_native_call, args=(q,), daemon=True)
    p.start()
    try:
        assert q.get(timeout=60) == 200000 * 8.0
    finally:
        p.join(timeout=5)
        if p.is_alive():
            p.terminate()
