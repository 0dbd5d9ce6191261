"""The backend tensor / sparse API that the reference exposes as ``dgl.backend`` (backend/backend.py): autograd
bookkeeping, sparse-matrix helpers, graph-level CSR algebra and the typed (heterogeneous) message-passing entry points -
each checked against a dense PyTorch computation."""
import importlib

import pytest
import torch

import dgl_b200 as dgl
from dgl_b200 import backend as F


def dense(g, w):
    u, v = g.edges()
    out = torch.zeros(g.num_src_nodes(), g.num_dst_nodes(), dtype=w.dtype)
    out.index_put_((u.long(), v.long()), w, accumulate=True)
    return out


def pair():
    torch.manual_seed(0)
    A = dgl.graph((torch.tensor([0, 1, 2, 2, 3]), torch.tensor([1, 2, 0, 1, 3])), num_nodes=4)
    B = dgl.graph((torch.tensor([0, 1, 1, 3]), torch.tensor([2, 0, 1, 0])), num_nodes=4)
    return A, torch.rand(5, dtype=torch.float64, requires_grad=True), B, torch.rand(4, dtype=torch.float64, requires_grad=True)


def test_csrmm_csrsum_csrmask_match_dense_and_differentiate():
    A, wa, B, wb = pair()
    C, wc = F.csrmm(A, wa, B, wb, 1)
    assert torch.allclose(dense(C, wc.detach()), dense(A, wa.detach()) @ dense(B, wb.detach()))
    (wc * torch.arange(1, len(wc) + 1)).sum().backward()
    a2, b2 = wa.detach().clone().requires_grad_(), wb.detach().clone().requires_grad_()
    prod = dense(A, a2) @ dense(B, b2)
    u, v = C.edges()
    (prod[u.long(), v.long()] * torch.arange(1, len(wc) + 1)).sum().backward()
    assert torch.allclose(wa.grad, a2.grad) and torch.allclose(wb.grad, b2.grad)
    S, ws = F.csrsum([A, B], [wa.detach(), wb.detach()])
    assert torch.allclose(dense(S, ws), dense(A, wa.detach()) + dense(B, wb.detach()))
    assert S.num_edges() == len({*zip(*map(torch.Tensor.tolist, A.edges()))} | {*zip(*map(torch.Tensor.tolist, B.edges()))})
    m = F.csrmask(A, wa.detach(), S)
    u, v = S.edges()
    assert torch.allclose(m, dense(A, wa.detach())[u.long(), v.long()])
    assert F.CSRMM.apply(A, wa, B, wb, 1)[1].shape == wc.shape and F.CSRMask.apply(A, wa, S).shape == m.shape
    bip = dgl.heterograph({("u", "e", "v"): (torch.tensor([0, 1]), torch.tensor([2, 0]))}, {"u": 2, "v": 3})
    bip2 = dgl.heterograph({("v", "e", "w"): (torch.tensor([2, 0, 0]), torch.tensor([1, 1, 0]))}, {"v": 3, "w": 2})
    P, wp = F.csrmm(bip, torch.ones(2), bip2, torch.ones(3), 2)
    assert P.num_src_nodes() == 2 and P.num_dst_nodes() == 2 and torch.allclose(
        dense(P, wp), dense(bip, torch.ones(2)) @ dense(bip2, torch.ones(3)))
    with pytest.raises(dgl.DGLError):
        F.csrsum([A], [])


def hetero():
    g = dgl.heterograph({("a", "r1", "b"): (torch.tensor([0, 1, 2, 2]), torch.tensor([0, 0, 1, 0])),
                         ("c", "r2", "b"): (torch.tensor([0, 1, 1]), torch.tensor([1, 2, 0])),
                         ("b", "r3", "a"): (torch.tensor([0, 2]), torch.tensor([2, 2]))},
                        {"a": 3, "b": 4, "c": 2})
    torch.manual_seed(1)
    feats = {nt: torch.randn(g.num_nodes(nt), 5) - 2.0 for nt in g.ntypes}          # mostly negative: a 0 fill would show
    return g, feats


@pytest.mark.parametrize("red", ["sum", "max", "min", "mean"])
def test_gspmm_hetero_reduces_over_all_relations_into_a_type(red):
    g, feats = hetero()
    out = F.gspmm_hetero(g, "copy_lhs", red, len(g.ntypes), *[feats[nt] for nt in g.ntypes])
    for ti, nt in enumerate(g.ntypes):
        incoming = [[] for _ in range(g.num_nodes(nt))]
        for s, e, d in g.canonical_etypes:
            if d != nt:
                continue
            u, v = g.edges(etype=(s, e, d))
            for a, b in zip(u.tolist(), v.tolist()):
                incoming[b].append(feats[s][a])
        if not any(incoming):
            assert out[ti] is None
            continue
        for node, msgs in enumerate(incoming):
            if not msgs:
                assert float(out[ti][node].abs().sum()) == 0.0
                continue
            stack = torch.stack(msgs)
            want = {"sum": stack.sum(0), "mean": stack.mean(0), "max": stack.max(0).values, "min": stack.min(0).values}[red]
            assert torch.allclose(out[ti][node], want, atol=1e-5), (red, nt, node)


def test_gspmm_hetero_edge_operands_and_grad():
    g, feats = hetero()
    ew = [torch.rand(g.num_edges(et), 1, requires_grad=True) for et in g.canonical_etypes]
    xs = [feats[nt].clone().requires_grad_() for nt in g.ntypes]
    out = F.gspmm_hetero(g, "mul", "sum", len(xs), *xs, *ew)
    bi = g.ntypes.index("b")
    want = torch.zeros(4, 5)
    for ri, (s, e, d) in enumerate(g.canonical_etypes):
        if d == "b":
            u, v = g.edges(etype=(s, e, d))
            want.index_add_(0, v.long(), xs[g.ntypes.index(s)].detach()[u.long()] * ew[ri].detach())
    assert torch.allclose(out[bi], want, atol=1e-5)
    out[bi].sum().backward()
    assert xs[g.ntypes.index("a")].grad is not None and ew[0].grad is not None and ew[1].grad is None


def test_gsddmm_and_edge_softmax_hetero():
    g, feats = hetero()
    xs = [feats[nt] for nt in g.ntypes]
    dots = F.gsddmm_hetero(g, "dot", len(xs), "u", "v", *xs, *xs)
    for ri, (s, e, d) in enumerate(g.canonical_etypes):
        u, v = g.edges(etype=(s, e, d))
        assert torch.allclose(dots[ri].squeeze(-1), (feats[s][u.long()] * feats[d][v.long()]).sum(-1), atol=1e-5)
    logits = [torch.randn(g.num_edges(et), 2) for et in g.canonical_etypes]
    sm = F.edge_softmax_hetero(g, None, "dst", *logits)
    for nt in g.ntypes:
        total = torch.zeros(g.num_nodes(nt), 2)
        zsum = torch.zeros(g.num_nodes(nt), 2)
        for ri, (s, e, d) in enumerate(g.canonical_etypes):
            if d == nt:
                v = g.edges(etype=(s, e, d))[1].long()
                total.index_add_(0, v, sm[ri])
                zsum.index_add_(0, v, torch.exp(logits[ri]))
        reached = zsum.sum(1) > 0
        assert torch.allclose(total[reached], torch.ones_like(total[reached]), atol=1e-5)
        for ri, (s, e, d) in enumerate(g.canonical_etypes):
            if d == nt:
                v = g.edges(etype=(s, e, d))[1].long()
                assert torch.allclose(sm[ri], torch.exp(logits[ri]) / zsum[v], atol=1e-5)
    by_src = F.edge_softmax_hetero(g, None, "src", *logits)
    u = g.edges(etype="r1")[0].long()
    assert torch.allclose(torch.zeros(3, 2).index_add_(0, u, by_src[0])[u.unique()], torch.ones(len(u.unique()), 2), atol=1e-5)


def test_autograd_and_tensor_helpers():
    x = torch.randn(4)
    F.attach_grad(x)
    assert x.requires_grad and F.is_no_grad(x) and F.is_recording()
    with F.record_grad():
        y = (x * x).sum()
    F.backward(y)
    assert torch.allclose(F.grad(x), 2 * x.detach()) and not F.is_no_grad(x)
    F.attach_grad(x)
    assert F.is_no_grad(x)                                     # re-attaching clears the gradient
    with F.no_grad():
        assert not F.is_recording()
    z = torch.arange(3.0, requires_grad=True)
    F.backward(z * 2, torch.ones(3))
    assert z.grad.tolist() == [2.0, 2.0, 2.0]
    F.sync()
    assert F.empty((2, 3), torch.float32, F.cpu()).shape == (2, 3) and F.cumsum(torch.ones(3), 0).tolist() == [1, 2, 3]
    assert F.floor_div(torch.tensor([7, -7]), torch.tensor([2, 2])).tolist() == [3, -4]
    assert torch.allclose(F.inverse(torch.eye(2) * 2), torch.eye(2) / 2) and F.get_preferred_sparse_format() == "coo"
    sp, perm = F.sparse_matrix(torch.tensor([1.0, 2.0, 3.0]), ("coo", torch.tensor([[0, 1, 2], [1, 2, 0]])), (3, 3))
    assert perm is None and sp.to_dense()[1, 2] == 2.0 and F.sparse_matrix_indices(sp)[0] == "coo"
    csr, _ = F.sparse_matrix(torch.tensor([1.0, 2.0, 3.0]), ("csr", (torch.tensor([1, 2, 0]), torch.tensor([0, 1, 2, 3]), None)), (3, 3))
    assert torch.equal(csr.to_dense(), sp.to_dense())
    with pytest.raises(TypeError):
        F.sparse_matrix(torch.ones(1), ("ell", None), (1, 1))
    t = torch.ones(2)
    assert F.zerocopy_to_dgl_ndarray(t) is t and F.zerocopy_from_dgl_ndarray(t) is t and F.zerocopy_to_dgl_ndarray_for_write(t) is t
    assert F.scatter_add(torch.ones(4, 2), torch.tensor([0, 0, 2, 1]), 3).tolist() == [[2, 2], [1, 1], [1, 1]]


def test_legacy_kernel_entry_points():
    A, wa, _, _ = pair()
    x, w = torch.randn(4, 3), wa.detach().float().unsqueeze(1)
    u, v = A.edges()
    want = torch.zeros(4, 3).index_add_(0, v.long(), x[u.long()] * w)
    assert torch.allclose(F.binary_reduce("sum", "mul", A, "src", "edge", x, w, 4), want, atol=1e-6)
    assert torch.allclose(F.binary_reduce("none", "add", A, "src", "dst", x, x, None), x[u.long()] + x[v.long()])
    deg = A.in_degrees().clamp(min=1).unsqueeze(1)
    assert torch.allclose(F.copy_reduce("mean", A, "src", x, 4), torch.zeros(4, 3).index_add_(0, v.long(), x[u.long()]) / deg, atol=1e-6)
    assert torch.allclose(F.copy_reduce("sum", A, "edge", w, 4), torch.zeros(4, 1).index_add_(0, v.long(), w))


def test_reference_backend_module_paths():
    for path, names in (("backend.backend", ["gspmm", "csrmm", "attach_grad", "sparse_matrix"]),
                        ("backend.pytorch.sparse", ["GSpMM", "GSDDMM", "EdgeSoftmax", "SegmentReduce", "CSRMM", "scatter_add"]),
                        ("backend.pytorch.tensor", ["zerocopy_from_dgl_ndarray", "cumsum", "data_type_dict"]),
                        ("backend.set_default_backend", ["set_default_backend"]),
                        ("ops.edge_softmax", ["edge_softmax"]), ("ops.gather_mm", ["gather_mm"]),
                        ("sparse.sddmm", ["sddmm", "bsddmm"]), ("sparse.matmul", ["spmm", "bspmm", "spspmm", "matmul"]),
                        ("sparse.softmax", ["softmax"]), ("transforms.to_block", ["to_block"]),
                        ("data.sbm", ["SBMMixtureDataset", "SBMMixture"]),
                        ("graphbolt.impl.gpu_cached_feature", ["GPUCachedFeature", "gpu_cached_feature"]),
                        ("graphbolt.impl.cpu_cached_feature", ["CPUCachedFeature", "cpu_cached_feature"])):
        mod = importlib.import_module("dgl_b200." + path)
        for n in names:
            assert getattr(mod, n) is not None, (path, n)
    # a module path that shares its name with a function of the package must not shadow the function
    import types
    g = dgl.graph((torch.tensor([0, 1]), torch.tensor([1, 1])))
    assert not isinstance(dgl.ops.edge_softmax, types.ModuleType)
    assert torch.allclose(dgl.ops.edge_softmax(g, torch.zeros(2)), torch.tensor([0.5, 0.5]))
    assert callable(dgl.to_block) and callable(dgl.transforms.to_block) and callable(dgl.sparse.softmax)
    assert callable(dgl.sparse.matmul) and callable(dgl.sparse.sddmm) and callable(dgl.ops.gather_mm)
    from dgl_b200.ops.edge_softmax import edge_softmax as via_path
    assert via_path is dgl.ops.edge_softmax
