"""tcgen05 fused SAGE layer kernel vs plain PyTorch fp32 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("Fp,Np,ncap,nvalid,fan,use_self", [(128, 256, 1000, 900, 10, True), (256, 256, 300, 300, 5, False),
                                                          (64, 64, 129, 1, 3, True), (128, 32, 4096, 4000, 10, True),
                                                          (128, 256, 30000, 29000, 10, True), (64, 128, 160000, 159990, 4, True)])
def test_sage_layer_tc(Fp, Np, ncap, nvalid, fan, use_self):
    from dgl_b200.ops.fused import sage_layer_tc
    torch.manual_seed(0)
    rows = 5000
    X = (torch.randn(rows, Fp, device=DEV) * 0.5).bfloat16()
    self_ids = torch.randint(0, rows, (ncap,), device=DEV, dtype=torch.int32) if use_self else None
    nbr = torch.randint(0, rows, (ncap, fan), device=DEV, dtype=torch.int32)
    cnt = torch.randint(0, fan + 1, (ncap,), device=DEV, dtype=torch.int32)
    n_dev = torch.tensor([nvalid], device=DEV, dtype=torch.int64)
    W = (torch.randn(2 * Fp, Np, device=DEV) * 0.05)
    Wt = W.t().contiguous().bfloat16()
    bias = torch.randn(Np, device=DEV)
    out, A = sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True, save_A=True)
    torch.cuda.synchronize()
    Xf = X.float()
    sid = self_ids.long() if use_self else torch.arange(ncap, device=DEV)
    mask = (torch.arange(fan, device=DEV).unsqueeze(0) < cnt.unsqueeze(1)).float().unsqueeze(-1)
    neigh = (Xf[nbr.long()] * mask).sum(1) / cnt.clamp(min=1).float().unsqueeze(1)
    Aref = torch.cat([Xf[sid], neigh], 1).bfloat16()
    assert torch.equal(A[:nvalid, :Fp], Aref[:nvalid, :Fp])
    assert (A[:nvalid].float() - Aref[:nvalid].float()).abs().max() < 2e-2
    ref = torch.relu(Aref.float() @ Wt.float().t() + bias)
    err = (out[:nvalid].float() - ref[:nvalid]).abs().max().item()
    scale = ref[:nvalid].abs().max().item()
    assert err < 2e-2 * max(scale, 1.0), (err, scale)


def test_sage_layer_tc_dropout():
    """Dropout in the epilogue: kept entries are the no-dropout values times 1/(1-p), the rest exactly 0."""
    from dgl_b200.ops.fused import sage_layer_tc
    torch.manual_seed(1)
    rows, Fp, Np, n, fan, p = 4000, 128, 256, 20000, 10, 0.25
    X = (torch.randn(rows, Fp, device=DEV) * 0.5).bfloat16()
    nbr = torch.randint(0, rows, (n, fan), device=DEV, dtype=torch.int32)
    cnt = torch.full((n,), fan, device=DEV, dtype=torch.int32)
    self_ids = torch.randint(0, rows, (n,), device=DEV, dtype=torch.int32)
    n_dev = torch.tensor([n], device=DEV, dtype=torch.int64)
    Wt = (torch.randn(Np, 2 * Fp, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(Np, device=DEV).abs() + 0.5  # mostly positive outputs so zeros come from dropout
    base, _ = sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True)
    step = torch.tensor([5], device=DEV, dtype=torch.int64)
    d1, _ = sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True, p_drop=p, seed=3, step_dev=step)
    d2, _ = sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True, p_drop=p, seed=3, step_dev=step)
    step += 1
    d3, _ = sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True, p_drop=p, seed=3, step_dev=step)
    torch.cuda.synchronize()
    assert torch.equal(d1, d2)          # counter-based: same (seed, step) -> same mask
    assert not torch.equal(d1, d3)      # new step -> new mask
    pos = base.float() > 0.05
    dropped = (d1.float() == 0) & pos
    frac = dropped.float().sum() / pos.float().sum()
    assert abs(frac.item() - p) < 0.01, frac
    kept = pos & ~dropped
    ratio = d1.float()[kept] / base.float()[kept]
    assert (ratio - 1.0 / (1.0 - p)).abs().max() < 0.02
    # masks must not be correlated along rows or columns
    m = dropped.float()
    assert abs(m.mean(0).std().item()) < 0.02 and abs(m.mean(1).std().item()) < 0.05
