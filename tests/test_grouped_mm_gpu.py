"""tcgen05 grouped GEMM (segment_mm / gather_mm forward + backward) against fp32 PyTorch references."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200.nn import functional_native as fnn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_segment_mm(a, b, seglen):
    outs, off = [], 0
    for r, n in enumerate(seglen.tolist()):
        outs.append(a[off:off + n].float() @ b[r].float())
        off += n
    return torch.cat(outs, 0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("K,M", [(64, 64), (128, 256), (72, 40), (256, 512), (16, 8), (200, 328)])
def test_segment_mm_tc_forward_backward(dtype, K, M):
    torch.manual_seed(0)
    seglen = torch.tensor([300, 0, 5, 1290, 128, 0, 77, 4100], device=DEV)
    R, N = seglen.shape[0], int(seglen.sum())
    a = (torch.randn(N, K, device=DEV) * 0.5).to(dtype).requires_grad_()
    b = (torch.randn(R, K, M, device=DEV) * 0.2).to(dtype).requires_grad_()
    assert fnn._tc_ok(a, b)
    out = dgl.ops.segment_mm(a, b, seglen)
    ref = _ref_segment_mm(a.detach(), b.detach(), seglen.cpu())
    tol = 2e-2 * max(1.0, K ** 0.5 * 0.1 * 4)
    assert out.dtype == dtype and (out.float() - ref).abs().max() < tol
    g = (torch.randn(N, M, device=DEV) * 0.5).to(dtype)
    out.backward(g)
    a32, b32 = a.detach().float().requires_grad_(), b.detach().float().requires_grad_()
    outs, off = [], 0
    for r, n in enumerate(seglen.tolist()):
        outs.append(a32[off:off + n] @ b32[r])
        off += n
    torch.cat(outs, 0).backward(g.float())
    assert (a.grad.float() - a32.grad).abs().max() < 2e-2 * max(1.0, M ** 0.5 * 0.2)
    rel = (b.grad.float() - b32.grad).abs().max() / b32.grad.abs().max().clamp_min(1e-6)
    assert rel < 2e-2, rel


def test_gather_mm_tc_unsorted():
    torch.manual_seed(1)
    N, K, M, R = 5000, 96, 64, 37
    a = (torch.randn(N, K, device=DEV) * 0.5).bfloat16().requires_grad_()
    b = (torch.randn(R, K, M, device=DEV) * 0.2).bfloat16().requires_grad_()
    idx = torch.randint(0, R, (N,), device=DEV)
    out = dgl.ops.gather_mm(a, b, idx_b=idx)
    ref = torch.bmm(a.detach().float().unsqueeze(1), b.detach().float()[idx]).squeeze(1)
    assert (out.float() - ref).abs().max() < 5e-2
    g = torch.randn(N, M, device=DEV).bfloat16()
    out.backward(g)
    a32, b32 = a.detach().float().requires_grad_(), b.detach().float().requires_grad_()
    torch.bmm(a32.unsqueeze(1), b32[idx]).squeeze(1).backward(g.float())
    assert (a.grad.float() - a32.grad).abs().max() < 5e-2
    rel = (b.grad.float() - b32.grad).abs().max() / b32.grad.abs().max()
    assert rel < 2e-2, rel


def test_typed_linear_uses_tc_path():
    lin = dgl.nn.TypedLinear(64, 32, 5).to(DEV).bfloat16()
    x = torch.randn(1000, 64, device=DEV).bfloat16()
    t = torch.sort(torch.randint(0, 5, (1000,), device=DEV))[0]
    y = lin(x, t, sorted_by_type=True)
    w = lin.get_weight().float()
    ref = torch.bmm(x.float().unsqueeze(1), w[t]).squeeze(1)
    assert (y.float() - ref).abs().max() < 5e-2
