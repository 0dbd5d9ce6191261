"""g-SpMM / g-SDDMM / edge_softmax / segment ops vs dense-torch oracles, forward and grads
(test strategy of /root/reference/tests/python/common/ops/test_ops.py:21-234)."""
import itertools

import pytest
import torch

import dgl_b200 as dgl
from dgl_b200 import ops
from graph_cases import get_cases


def _edge_list(g):
    u, v = g.edges()
    return u.long(), v.long()


def oracle_spmm(g, op, red, x, w):
    u, v = _edge_list(g)
    n_dst = g.num_dst_nodes()
    if op == "copy_lhs":
        m = x[u]
    elif op == "copy_rhs":
        m = w
    else:
        a, b = x[u], w
        m = {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b}[op]
    shape = (n_dst,) + tuple(m.shape[1:])
    if red in ("sum", "mean"):
        out = torch.zeros(shape, dtype=m.dtype).index_add_(0, v, m)
        if red == "mean":
            deg = torch.bincount(v, minlength=n_dst).clamp(min=1).to(m.dtype)
            out = out / deg.view(-1, *([1] * (m.dim() - 1)))
        return out
    fill = float("-inf") if red == "max" else float("inf")
    out = torch.full(shape, fill, dtype=m.dtype)
    idx = v.view(-1, *([1] * (m.dim() - 1))).expand_as(m)
    out = out.scatter_reduce(0, idx, m, reduce="amax" if red == "max" else "amin", include_self=True)
    return torch.where(torch.isinf(out), torch.zeros_like(out), out)


SHAPES = [((5,), (5,)), ((4, 1), (4, 3)), ((1, 3), (4, 1)), ((3,), (1,)), ((2, 3, 1), (1, 3, 4)), ((1,), (3,))]


@pytest.mark.parametrize("case", get_cases(["homo", "bipartite"]), ids=lambda f: f.__name__)
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "copy_lhs", "copy_rhs"])
@pytest.mark.parametrize("red", ["sum", "max", "min", "mean"])
@pytest.mark.parametrize("shp", SHAPES[:4], ids=str)
def test_gspmm(case, op, red, shp, idtype):
    g = case(idtype)
    torch.manual_seed(0)
    nu, ne = g.num_src_nodes(), g.num_edges()
    x = (torch.rand((nu,) + shp[0], dtype=torch.float64) + 0.5).requires_grad_()
    w = (torch.rand((ne,) + shp[1], dtype=torch.float64) + 0.5).requires_grad_()
    out = ops.gspmm(g, op, red, x, w)
    ref = oracle_spmm(g, op, red, x, w)
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, atol=1e-9)
    gout = torch.rand_like(ref)
    need = [t for t, used in ((x, op != "copy_rhs"), (w, op != "copy_lhs")) if used]
    g1 = torch.autograd.grad(out, need, gout, retain_graph=True)
    g2 = torch.autograd.grad(ref, need, gout)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, atol=1e-8), (a - b).abs().max()


def oracle_sddmm(g, op, lhs, rhs, lt, rt):
    u, v = _edge_list(g)
    sel = {"u": lambda t: t[u], "v": lambda t: t[v], "e": lambda t: t}
    a = sel[lt](lhs) if lhs is not None else None
    b = sel[rt](rhs) if rhs is not None else None
    if op == "dot":
        return (a * b).sum(-1, keepdim=True)
    if op == "copy_lhs":
        return a
    if op == "copy_rhs":
        return b
    return {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b}[op]


@pytest.mark.parametrize("case", get_cases(["homo", "bipartite"]), ids=lambda f: f.__name__)
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "dot", "copy_lhs", "copy_rhs"])
@pytest.mark.parametrize("lt,rt", [("u", "v"), ("u", "e"), ("e", "v"), ("v", "u"), ("v", "e"), ("e", "u")])
@pytest.mark.parametrize("shp", [((5,), (5,)), ((4, 1), (4, 3)), ((2, 3), (2, 3))], ids=str)
def test_gsddmm(case, op, lt, rt, shp, idtype):
    g = case(idtype)
    torch.manual_seed(1)
    n = {"u": g.num_src_nodes(), "v": g.num_dst_nodes(), "e": g.num_edges()}
    s0, s1 = shp
    if op == "dot":
        s0 = s0[:-1] + (3,)
        s1 = s1[:-1] + (3,)
    x = (torch.rand((n[lt],) + s0, dtype=torch.float64) + 0.5).requires_grad_()
    y = (torch.rand((n[rt],) + s1, dtype=torch.float64) + 0.5).requires_grad_()
    out = ops.gsddmm(g, op, x, y, lt, rt)
    ref = oracle_sddmm(g, op, x, y, lt, rt)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    assert torch.allclose(out, ref, atol=1e-9)
    gout = torch.rand_like(ref)
    need = [t for t, used in ((x, op != "copy_rhs"), (y, op != "copy_lhs")) if used]
    g1 = torch.autograd.grad(out, need, gout, retain_graph=True)
    g2 = torch.autograd.grad(ref, need, gout)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, atol=1e-8), (a - b).abs().max()


@pytest.mark.parametrize("case", get_cases(["homo", "bipartite"]), ids=lambda f: f.__name__)
@pytest.mark.parametrize("norm_by", ["dst", "src"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_edge_softmax(case, norm_by, dtype, idtype):
    g = case(idtype)
    torch.manual_seed(2)
    s = torch.randn(g.num_edges(), 4, 1, dtype=dtype, requires_grad=True)
    out = ops.edge_softmax(g, s, norm_by=norm_by)
    u, v = _edge_list(g)
    key = v if norm_by == "dst" else u
    n = g.num_dst_nodes() if norm_by == "dst" else g.num_src_nodes()
    mx = torch.full((n, 4, 1), float("-inf"), dtype=dtype).scatter_reduce(
        0, key.view(-1, 1, 1).expand_as(s), s, reduce="amax")
    e = torch.exp(s - mx[key])
    den = torch.zeros((n, 4, 1), dtype=dtype).index_add_(0, key, e)
    ref = e / den[key]
    assert torch.allclose(out, ref, atol=1e-6)
    go = torch.rand_like(ref)
    (g1,) = torch.autograd.grad(out, s, go, retain_graph=True)
    (g2,) = torch.autograd.grad(ref, s, go)
    assert torch.allclose(g1, g2, atol=1e-5)


@pytest.mark.parametrize("red", ["sum", "max", "min", "mean"])
def test_segment_reduce(red):
    torch.manual_seed(3)
    seglen = torch.tensor([3, 0, 4, 1, 0])
    x = torch.randn(8, 5, dtype=torch.float64, requires_grad=True)
    out = ops.segment_reduce(seglen, x, red)
    refs = []
    off = 0
    for n in seglen.tolist():
        seg = x[off:off + n]
        off += n
        if n == 0:
            refs.append(torch.zeros(5, dtype=torch.float64))
        else:
            refs.append({"sum": seg.sum(0), "mean": seg.mean(0), "max": seg.max(0)[0], "min": seg.min(0)[0]}[red])
    ref = torch.stack(refs)
    assert torch.allclose(out, ref)
    go = torch.rand_like(ref)
    (g1,) = torch.autograd.grad(out, x, go, retain_graph=True)
    (g2,) = torch.autograd.grad(ref, x, go)
    assert torch.allclose(g1, g2)


def test_segment_softmax_and_mm():
    seglen = torch.tensor([2, 3, 1])
    x = torch.randn(6, 3)
    out = ops.segment_softmax(seglen, x)
    assert torch.allclose(out[:2], torch.softmax(x[:2], 0), atol=1e-6)
    a = torch.randn(6, 4, requires_grad=True)
    b = torch.randn(3, 4, 5, requires_grad=True)
    y = ops.segment_mm(a, b, seglen)
    ref = torch.cat([a[:2] @ b[0], a[2:5] @ b[1], a[5:] @ b[2]])
    assert torch.allclose(y, ref, atol=1e-6)
    idx = torch.tensor([0, 2, 1, 1, 0, 2])
    z = ops.gather_mm(a, b, idx_b=idx)
    ref2 = torch.stack([a[i] @ b[idx[i]] for i in range(6)])
    assert torch.allclose(z, ref2, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_spmm(dtype):
    g = get_cases(["random"])[0]()
    x = torch.rand(g.num_nodes(), 16).to(dtype)
    out = ops.copy_u_sum(g, x)
    ref = oracle_spmm(g, "copy_lhs", "sum", x.float(), None)
    assert torch.allclose(out.float(), ref, rtol=2e-2, atol=2e-2)
