"""tcgen05 backward kernels (csrc/cuda/sage_bwd_tc.cu) and the fp32 (3xTF32) variant of the fused layer kernel
against plain PyTorch references (fp64 for the fp32 kernels, so the fp32-level accuracy claim is actually checked)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ell(ncap, fan, rows, seed=0):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    nbr = torch.randint(0, rows, (ncap, fan), device=DEV, dtype=torch.int32, generator=g)
    cnt = torch.randint(0, fan + 1, (ncap,), device=DEV, dtype=torch.int32, generator=g)
    return nbr, cnt


@pytest.mark.parametrize("Fp,Np,ncap,nvalid,fan,use_self", [(128, 256, 1000, 900, 10, True), (256, 256, 300, 300, 5, False),
                                                          (64, 64, 129, 1, 3, True), (32, 128, 4096, 4000, 10, True),
                                                          (128, 256, 30000, 29000, 10, True)])
def test_sage_layer_tc_fp32(Fp, Np, ncap, nvalid, fan, use_self):
    """fp32 operands: the 3xTF32 split must reproduce an fp64 reference to fp32-level accuracy (not TF32-level)."""
    from dgl_b200.ops.fused import sage_layer_tc
    torch.manual_seed(0)
    rows = 5000
    X = torch.randn(rows, Fp, device=DEV) * 0.5
    self_ids = torch.randint(0, rows, (ncap,), device=DEV, dtype=torch.int32) if use_self else None
    nbr, cnt = _ell(ncap, fan, rows)
    n_dev = torch.tensor([nvalid], device=DEV, dtype=torch.int64)
    Wt = torch.randn(Np, 2 * Fp, device=DEV) * 0.05
    bias = torch.randn(Np, device=DEV)
    out, A = sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True, save_A=True)
    torch.cuda.synchronize()
    Xd = X.double()
    sid = self_ids.long() if use_self else torch.arange(ncap, device=DEV)
    mask = (torch.arange(fan, device=DEV).unsqueeze(0) < cnt.unsqueeze(1)).double().unsqueeze(-1)
    neigh = (Xd[nbr.long()] * mask).sum(1) / cnt.clamp(min=1).double().unsqueeze(1)
    Aref = torch.cat([Xd[sid], neigh], 1)
    assert torch.equal(A[:nvalid, :Fp], X[sid][:nvalid])
    assert (A[:nvalid].double() - Aref[:nvalid]).abs().max() < 1e-6
    ref = torch.relu(Aref @ Wt.double().t() + bias.double())
    err = (out[:nvalid].double() - ref[:nvalid]).abs().max().item()
    scale = ref[:nvalid].abs().max().item()
    # TF32 alone would sit around 1e-3 * scale here; the split products have to land at fp32 rounding level
    assert err < 2e-5 * max(scale, 1.0), (err, scale)


def _dw_case(dt, n_cap, n, CP, CQ, transform, transposed, nbs=0):
    from dgl_b200.ops.fused import sage_bwd_dw
    torch.manual_seed(1)
    n_dev = torch.tensor([n], device=DEV, dtype=torch.int64)
    Q = (torch.randn(n_cap, CQ, device=DEV) * 0.5).to(dt)
    out_shape = (CQ, CP) if transposed else (CP, CQ)
    out = torch.zeros(out_shape, device=DEV)
    if transform:
        dH = torch.randn(n_cap, CP, device=DEV)
        H = torch.relu(torch.randn(n_cap, CP, device=DEV)).to(dt)
        gb = torch.zeros(CP, device=DEV)
        dz = torch.zeros(n_cap, CP, device=DEV, dtype=dt)
        dz_lo = torch.zeros(n_cap, CP, device=DEV) if dt == torch.float32 else None
        sage_bwd_dw(out, Q, n_dev, dH=dH, Hmask=H, scale=2.0, transposed_out=transposed, gb=gb, dz_out=dz, dz_out_lo=dz_lo, nbs=nbs)
        dZ = torch.where(H.float() > 0, dH * 2.0, torch.zeros_like(dH))
        dZq = dZ.to(dt)
        torch.cuda.synchronize()
        got = dz.float() + (dz_lo if dz_lo is not None else 0)
        assert torch.equal(got[:n], dZq[:n].float())
        assert (gb.double() - dZ[:n].double().sum(0)).abs().max() < 1e-3 * max(1.0, dZ[:n].abs().sum(0).max().item())
        P = dZq
    else:
        P = (torch.randn(n_cap, CP, device=DEV) * 0.5).to(dt)
        sage_bwd_dw(out, Q, n_dev, P=P, transposed_out=transposed, nbs=nbs)
        torch.cuda.synchronize()
    ref = P[:n].double().t() @ Q[:n].double()
    if transposed:
        ref = ref.t()
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-5 if dt == torch.float32 else 1e-3   # bf16 inputs are exact products: only fp32 accumulation order differs
    assert err < tol * max(scale, 1.0), (err, scale)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n_cap,n,CP,CQ,transform,transposed,nbs", [
    (1000, 937, 256, 256, True, False, 0),      # hidden layer, dZ built on the fly
    (30000, 29111, 256, 256, True, False, 128),  # split into 128-wide N blocks (4 quadrants)
    (5000, 5000, 128, 512, True, False, 0),
    (1024, 1000, 512, 64, False, True, 0),      # output layer: M side = A columns, transposed store
    (200, 3, 256, 128, False, False, 0),        # a single live chunk
    (70000, 69999, 256, 256, True, False, 0),   # many chunks per CTA
])
def test_sage_bwd_dw(dt, n_cap, n, CP, CQ, transform, transposed, nbs):
    _dw_case(dt, n_cap, n, CP, CQ, transform, transposed, nbs)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n_cap,n,Np,F,fan,rows_below,cb", [
    (1024, 1000, 64, 256, 10, 9000, 0), (11000, 10500, 256, 256, 10, 90000, 0), (11000, 10500, 256, 256, 10, 90000, 256),
    (5000, 4999, 256, 256, 10, 40000, 128), (300, 300, 128, 128, 5, 2000, 0), (129, 1, 256, 64, 3, 500, 0)])
def test_sage_bwd_da(dt, n_cap, n, Np, F, fan, rows_below, cb):
    from dgl_b200.ops.fused import sage_bwd_da
    torch.manual_seed(2)
    nbr, cnt = _ell(n_cap, fan, rows_below, seed=3)
    n_dev = torch.tensor([n], device=DEV, dtype=torch.int64)
    dz = (torch.randn(n_cap, Np, device=DEV) * 0.5).to(dt)
    Wt = (torch.randn(Np, 2 * F, device=DEV) * 0.1).to(dt)
    dH = torch.zeros(rows_below, F, device=DEV)
    sage_bwd_da(dz, Wt, nbr, cnt, n_dev, dH, cb=cb)
    torch.cuda.synchronize()
    dA = dz[:n].double() @ Wt.double()
    ref = torch.zeros(rows_below, F, device=DEV, dtype=torch.float64)
    ref[:n] += dA[:, :F]
    w = 1.0 / cnt[:n].clamp(min=1).double()
    contrib = dA[:, F:] * w.unsqueeze(1)
    mask = torch.arange(fan, device=DEV).unsqueeze(0) < cnt[:n].unsqueeze(1)
    idx = nbr[:n].long()[mask]
    src = contrib.unsqueeze(1).expand(-1, fan, -1)[mask]
    ref.index_add_(0, idx, src)
    err = (dH.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-5 if dt == torch.float32 else 1e-3
    assert err < tol * max(scale, 1.0), (err, scale)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_fused_sageconv_autograd_tc(dt):
    """dgl.nn.SAGEConv(mean) on a block through the tcgen05 forward + backward kernels vs the dense formulation."""
    import dgl_b200 as dgl
    from dgl_b200.nn import fused_sage
    torch.manual_seed(4)
    n_src, n_dst, F, O, fan = 3000, 700, 128, 200, 6
    nbr, cnt = _ell(n_dst, fan, n_src, seed=5)
    cnt = cnt.clamp(min=1)
    mask = torch.arange(fan, device=DEV).unsqueeze(0) < cnt.unsqueeze(1)
    x = (torch.randn(n_src, F, device=DEV) * 0.5).to(dt).requires_grad_(True)
    ws = (torch.randn(O, F, device=DEV) * 0.1).to(dt).requires_grad_(True)
    wn = (torch.randn(O, F, device=DEV) * 0.1).to(dt).requires_grad_(True)
    b = torch.randn(O, device=DEV, requires_grad=True)
    out = fused_sage._FusedSAGETC.apply(x, ws, wn, b, nbr, cnt, fan, True, None)
    gout = torch.randn_like(out)
    out.backward(gout)
    torch.cuda.synchronize()
    xd, wsd, wnd, bd = (t.detach().double().requires_grad_(True) for t in (x, ws, wn, b))
    neigh = (xd[nbr.long()] * mask.unsqueeze(-1)).sum(1) / cnt.double().unsqueeze(1)
    z = xd[:n_dst] @ wsd.t() + neigh @ wnd.t() + bd
    # the kernel's ReLU mask is taken on its own (rounded) output: the reference backward uses the SAME mask, otherwise
    # the handful of pre-activations within rounding distance of 0 would dominate the comparison
    (z * (out.detach() > 0)).backward(gout.double())
    tol = 3e-5 if dt == torch.float32 else 3e-2

    def check(name, a, r):
        err, scale = (a.double() - r).abs().max().item(), max(1.0, r.abs().max().item())
        assert err <= tol * scale, (name, err, scale)

    check("out", out, torch.relu(z.detach()))
    for name, a, r in (("dx", x.grad, xd.grad), ("dw_self", ws.grad, wsd.grad), ("dw_neigh", wn.grad, wnd.grad),
                       ("db", b.grad, bd.grad)):
        check(name, a, r)
