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
