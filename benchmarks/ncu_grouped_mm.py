"""Small driver for `ncu --set full -k regex:seg_mm`: one forward+backward of segment_mm on an R-GCN-like shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgl_b200 as dgl

N, R, K, M = 2_000_000, 64, 128, 128
seglen = torch.bincount(torch.randint(0, R, (N,)), minlength=R).cuda()
a = torch.randn(N, K, device="cuda").bfloat16().requires_grad_()
b = (torch.randn(R, K, M, device="cuda") * 0.1).bfloat16().requires_grad_()
go = torch.randn(N, M, device="cuda").bfloat16()
for _ in range(3):
    a.grad = b.grad = None
    dgl.ops.segment_mm(a, b, seglen).backward(go)
torch.cuda.synchronize()
