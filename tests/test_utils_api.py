"""``dgl.utils`` helper surface (reference python/dgl/utils/{internal,checks,data,filter,exception}.py;
reference tests: tests/python/common/utils/test_utils.py, test_filter.py, tests/python/common/test_index.py)."""
import pickle

import networkx as nx
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import dgl_b200 as dgl
from dgl_b200 import utils as U


def test_index_views_and_algebra():
    idx = U.toindex([1, 4, 2, 0])
    assert len(idx) == 4 and list(idx) == [1, 4, 2, 0] and idx[1] == 4
    assert idx.tonumpy().tolist() == [1, 4, 2, 0]
    assert idx.tousertensor().dtype == torch.int64
    assert U.toindex(idx) is idx
    sl = U.toindex(slice(2, 6))
    assert sl.is_slice(2, 6) and sl.slice_data() == slice(2, 6) and list(sl) == [2, 3, 4, 5]
    assert list(sl.get_items(U.toindex(slice(1, 3)))) == [3, 4]
    assert sl.get_items(U.toindex(slice(1, 3))).is_slice(3, 5)
    assert list(idx.get_items(U.toindex([0, 0, 3]))) == [1, 1, 0]
    assert list(idx.set_items(U.toindex([0]), 9)) == [9, 4, 2, 0]
    assert list(idx.append_zeros(2)) == [1, 4, 2, 0, 0, 0]
    assert list(idx.nonzero()) == [0, 1, 2] and idx.has_nonzero()
    assert not U.zero_index(3).has_nonzero()
    for obj in (idx, sl):
        back = pickle.loads(pickle.dumps(obj))
        assert list(back) == list(obj) and back.slice_data() == obj.slice_data()
    with pytest.raises(dgl.DGLError):
        U.toindex(torch.tensor([1.0, 2.0]))
    with pytest.raises(U.InconsistentDtypeException):
        U.toindex(torch.tensor([1, 2], dtype=torch.int32), "int64")
    assert list(U.set_diff(U.toindex([5, 1, 3, 3]), U.toindex([3]))) == [1, 5]


def test_dict_views():
    lazy = U.LazyDict(lambda k: k * 2, ["a", "b"])
    assert lazy["a"] == "aa" and "b" in lazy and len(lazy) == 2 and list(lazy) == ["a", "b"]
    with pytest.raises(KeyError):
        lazy["z"]
    hy = U.HybridDict({"a": 1}, {"a": 2, "b": 3})
    assert hy["a"] == 1 and hy["b"] == 3 and set(hy) == {"a", "b"} and len(hy) == 2
    ro = U.ReadOnlyDict({"x": 1})
    assert ro["x"] == 1 and "x" in ro and len(ro) == 1
    with pytest.raises(TypeError):
        ro["y"] = 2
    fd = U.FlattenedDict({"a": [1, 3], "b": [2, 5, 8], "c": [7]})
    assert list(fd) == [("a", 1), ("a", 3), ("b", 2), ("b", 5), ("b", 8), ("c", 7)]
    assert fd[2] == ("b", 2) and fd[5] == ("c", 7) and len(fd) == 6
    assert U.maybe_flatten_dict([1, 2]) == [1, 2] and isinstance(U.maybe_flatten_dict({"a": [1]}), U.FlattenedDict)
    assert dict(U.group_as_dict([("a", 1), ("b", 2), ("a", 3)])) == {"a": [1, 3], "b": [2]}
    assert U.is_dict_like({}) and not U.is_dict_like([])
    assert U.is_listlike((1,)) and not U.is_listlike("ab") and U.is_iterable("ab")


def test_relabel_family():
    x = torch.tensor([1, 5, 3, 6, 5])
    n2o, o2n = U.relabel(x)
    assert n2o.tolist() == [1, 3, 5, 6] and o2n.tolist() == [0, 0, 0, 1, 0, 2, 3]
    assert torch.equal(n2o[o2n[x]], x)
    n2o, o2n = U.build_relabel_map(U.toindex([1, 5, 3, 6]))
    assert n2o.tolist() == [1, 3, 5, 6] and o2n[[1, 3, 5, 6]].tolist() == [0, 1, 2, 3]
    assert U.build_relabel_dict([7, 3]) == {7: 0, 3: 1}
    assert U.compensate(torch.tensor([0, 2, 4]), torch.tensor([0, 1, 2, 4, 5])).tolist() == [1, 5]
    uniq, inv, remapped = U.make_invmap([10, 30, 10, 20])
    assert list(uniq) == [10, 20, 30] and remapped.tolist() == [0, 2, 0, 1] and inv[30] == 2
    assert U.to_nbits_int(torch.tensor([1]), 32).dtype == torch.int32
    d = U.reorder({"h": torch.arange(4.0)}, U.toindex([3, 0]))
    assert d["h"].tolist() == [3.0, 0.0]
    assert list(U.reorder_index(U.toindex([5, 6, 7]), U.toindex([2, 0]))) == [7, 5]


def test_decorators_and_apply():
    class Obj:
        def __init__(self):
            self._cache, self.calls, self.fixed = {}, 0, False

        @U.cached_member("_cache", "sq")
        def square(self, x):
            self.calls += 1
            return x * x

        def fix(self):
            self.fixed = True

        @U.retry_method_with_fix(fix)
        def fragile(self):
            if not self.fixed:
                raise RuntimeError("not yet")
            return "ok"

    o = Obj()
    assert o.square(3) == 9 and o.square(3) == 9 and o.calls == 1 and o.square(4) == 16 and o.calls == 2
    assert o.fragile() == "ok" and o.fixed
    gen = U.CtxCachedObject(lambda ctx: object())
    assert gen("cpu") is gen("cpu") and gen("cpu") is not gen("cuda")
    assert U.apply_each({"a": 1, "b": 2}, lambda v, k: v + k, 10) == {"a": 11, "b": 12}
    assert U.apply_each([1, 2], lambda v: -v) == [-1, -2] and U.apply_each(3, lambda v: -v) == -3
    out = U.recursive_apply_pair({"a": (1, [2, 3])}, {"a": (10, [20, 30])}, lambda x, y: x + y)
    assert out == {"a": (11, [22, 33])}
    alias = U.alias_func(dgl.add_self_loop)
    assert "Alias of" in alias.__doc__ and alias(dgl.graph(([0], [1]))).num_edges() == 3


def test_threads_and_numa_queries():
    n = U.get_num_threads()
    assert n >= 1
    U.set_num_threads(2)
    assert U.get_num_threads() == 2 and torch.get_num_threads() == 2
    U.set_num_threads(n)
    assert isinstance(U.get_numa_nodes_cores(), dict)
    assert U.to_dgl_context("cpu") == torch.device("cpu")


def test_checks():
    g = dgl.graph(([0, 1, 2], [1, 2, 0]), idtype=torch.int32)
    assert U.prepare_tensor_or_dict(g, [0, 1], "x").dtype == torch.int32
    assert U.prepare_tensor_or_dict(g, {"_N": [0]}, "x")["_N"].dtype == torch.int32
    assert U.parse_edges_arg_to_eid(g, (torch.tensor([1]), torch.tensor([2])), 0).tolist() == [1]
    assert U.parse_edges_arg_to_eid(g, [2, 0], 0).tolist() == [2, 0]
    g64 = dgl.graph(([0], [1]))
    U.check_all_same_idtype([g, g], "gs")
    with pytest.raises(dgl.DGLError):
        U.check_all_same_idtype([g, g64], "gs")
    U.check_all_same_device([g, g64], "gs")
    assert U.check_device(torch.zeros(1), "cpu") and U.check_device({"a": torch.zeros(1)}, torch.device("cpu"))
    U.check_all_same_schema([{"h": 1}, {"h": 1}], "ndata")
    with pytest.raises(dgl.DGLError):
        U.check_all_same_schema([{"h": 1}, {"h": 2}], "ndata")
    U.check_all_same_schema_for_keys([{"h": 1, "x": 0}, {"h": 1, "x": 5}], ["h"], "ndata")
    with pytest.raises(dgl.DGLError):
        U.check_all_same_schema_for_keys([{"h": 1}, {"x": 1}], ["h"], "ndata")
    with pytest.raises(dgl.DGLError):
        U.check_all_same_schema_for_keys([{"h": 1}, {"h": 2}], ["h"], "ndata")
    U.check_valid_idtype(None)
    U.check_valid_idtype(torch.int32)
    with pytest.raises(dgl.DGLError):
        U.check_valid_idtype(torch.float32)
    assert U.is_sorted_srcdst(torch.tensor([0, 0, 1, 2]), torch.tensor([1, 3, 0, 0])) == (True, True)
    assert U.is_sorted_srcdst(torch.tensor([0, 0, 1, 2]), torch.tensor([3, 1, 0, 0])) == (True, False)
    assert U.is_sorted_srcdst(torch.tensor([1, 0]), torch.tensor([0, 0])) == (False, False)


def test_graphdata2tensors_inputs():
    (fmt, (u, v)), ns, nd = U.graphdata2tensors([(0, 1), (2, 3)])
    assert fmt == "coo" and u.tolist() == [0, 2] and v.tolist() == [1, 3] and (ns, nd) == (4, 4)
    (_, (u, v)), ns, nd = U.graphdata2tensors((torch.tensor([0, 1], dtype=torch.int32), torch.tensor([4, 0], dtype=torch.int32)),
                                              bipartite=True)
    assert u.dtype == torch.int32 and (ns, nd) == (2, 5)
    (fmt, arrays), ns, nd = U.graphdata2tensors(("csr", ([0, 1, 2], [1, 0], [])))
    assert fmt == "csr" and arrays[0].dtype == torch.int64 and (ns, nd) == (2, 2)
    m = sp.random(5, 7, density=0.3, format="csr", random_state=0)
    (fmt, arrays), ns, nd = U.graphdata2tensors(m, idtype=torch.int32, bipartite=True)
    assert fmt == "csr" and arrays[0].tolist() == m.indptr.tolist() and (ns, nd) == (5, 7)
    (fmt, arrays), _, _ = U.graphdata2tensors(m.tocoo())
    assert fmt == "coo" and arrays[0].tolist() == m.tocoo().row.tolist()
    G = nx.path_graph(4)
    (_, (u, v)), ns, nd = U.graphdata2tensors(G)
    assert (ns, nd) == (4, 4) and u.shape[0] == 6
    D = nx.DiGraph()
    D.add_edge("b", "a", id=1)
    D.add_edge("a", "c", id=0)
    u, v = U.networkx2tensor(D, torch.int64, edge_id_attr_name="id")
    assert u.tolist() == [0, 1] and v.tolist() == [2, 0]
    B = nx.Graph()
    B.add_nodes_from(["x", "y"], bipartite=0)
    B.add_nodes_from([10, 11, 12], bipartite=1)
    B.add_edges_from([("x", 10), ("y", 12)])
    (_, (u, v)), ns, nd = U.graphdata2tensors(B, bipartite=True, top_map={"x": 0, "y": 1}, bottom_map={10: 0, 11: 1, 12: 2})
    assert (ns, nd) == (2, 3) and sorted(zip(u.tolist(), v.tolist())) == [(0, 0), (1, 2)]
    with pytest.raises(dgl.DGLError):
        U.networkxbipartite2tensors(B, torch.int64, {"x": 0}, {10: 0, 12: 1})
    assert U.infer_num_nodes(object()) is None
    with pytest.raises(dgl.DGLError):
        U.graphdata2tensors(3.5)
    moved = U.to_device({"a": torch.zeros(1)}, "cpu")
    assert moved["a"].device.type == "cpu"


def test_filter_included_excluded():
    f = U.Filter(torch.tensor([3, 2, 9]))
    test = torch.tensor([0, 2, 8, 9])
    assert f.find_included_indices(test).tolist() == [1, 3]
    assert f.find_excluded_indices(test).tolist() == [0, 2]
    empty = U.Filter(torch.tensor([], dtype=torch.int64))
    assert empty.find_included_indices(test).numel() == 0 and empty.find_excluded_indices(test).tolist() == [0, 1, 2, 3]
    big = torch.randperm(100000)[:5000]
    q = torch.randint(0, 100000, (20000,))
    inc = U.Filter(big).find_included_indices(q)
    assert torch.equal(inc, torch.nonzero(torch.isin(q, big)).view(-1))


def test_exception_wrapper_roundtrip():
    try:
        {}["missing"]
    except KeyError:
        w = U.ExceptionWrapper(where="in worker 3")
    w = pickle.loads(pickle.dumps(w))
    with pytest.raises(KeyError) as ei:
        w.reraise()
    assert "in worker 3" in str(ei.value) and "missing" in str(ei.value)
    try:
        raise ValueError("boom")
    except ValueError:
        w = U.ExceptionWrapper()
    with pytest.raises(ValueError, match="boom"):
        w.reraise()
    assert repr(U.KeyErrorMessage("a\nb")) == "a\nb"


def test_shared_mem_aliases_and_block_subframes():
    name = "dglb200_test_utils_arr"
    a = U.create_shared_mem_array(name, (3, 2), torch.float32)
    a.fill_(4.0)
    b = U.get_shared_mem_array(name, (3, 2), torch.float32)
    assert torch.equal(a, b)
    from dgl_b200.utils.shared_mem import unlink
    unlink(name)
    g = dgl.graph(([0, 1, 2], [1, 2, 3]))
    g.ndata["h"] = torch.arange(4.0)
    frames = U.extract_node_subframes_for_block(g, [torch.tensor([3, 1])], [torch.tensor([1])])
    assert frames[0]["h"].tolist() == [3.0, 1.0] and frames[1]["h"].tolist() == [1.0]
    assert frames[0][dgl.NID].tolist() == [3, 1]


def test_container_and_id_helpers():
    """``dgl.utils.internal`` re-exports the container / relabelling helpers of utils/containers.py."""
    import torch
    from dgl_b200.utils import internal as I
    u, m = I.relabel(torch.tensor([1, 5, 3, 6]))
    assert u.tolist() == [1, 3, 5, 6] and m[[1, 3, 5, 6]].tolist() == [0, 1, 2, 3]
    assert I.build_relabel_map(torch.tensor([1, 5, 3, 6]))[1][[1, 3, 5, 6]].tolist() == [0, 1, 2, 3]
    assert I.set_diff(torch.tensor([1, 2, 3, 4]), torch.tensor([2, 4])).tousertensor().tolist() == [1, 3]
    assert I.compensate(torch.tensor([0, 2]), torch.arange(4)).tolist() == [1, 3]
    f = I.FlattenedDict({"a": [10, 11], "b": [20]})
    assert len(f) == 3 and f[2] == ("b", 20) and f[0] == ("a", 10)
    h = I.HybridDict({"x": 1}, {"x": 2, "y": 3})
    assert h["x"] == 1 and h["y"] == 3
    assert dict(I.group_as_dict([("a", 1), ("b", 2), ("a", 3)])) == {"a": [1, 3], "b": [2]}
    assert I.recursive_apply_pair([1, {"k": 2}], [10, {"k": 20}], lambda a, b: a + b) == [11, {"k": 22}]
    _, inv, rm = I.make_invmap([7, 3, 7])
    assert {int(k): v for k, v in inv.items()} == {3: 0, 7: 1} and list(rm) == [1, 0, 1]
    assert I.to_nbits_int(torch.tensor([1]), 32).dtype == torch.int32
    assert I.toindex([1, 2]).tousertensor().tolist() == [1, 2] and I.get_num_threads() >= 1
