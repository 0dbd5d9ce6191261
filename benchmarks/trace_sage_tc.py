"""Cycle-stamp trace of CTA 0 of the fused SAGE layer kernel (where does a tile's time go?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgl_b200.ops.fused import sage_layer_tc

DEV = "cuda"
for (rows, n, Fp, fan, cz, pd) in [(2_400_000, 59520, 128, 10, False, 0.5), (2_400_000, 59520, 128, 10, True, 0.0), (2_400_000, 8704, 256, 10, False, 0.5)]:
    torch.manual_seed(0)
    X = (torch.randn(rows, Fp, device=DEV) * 0.5).bfloat16()
    self_ids = torch.randint(0, rows, (n,), device=DEV, dtype=torch.int32)
    nbr = torch.randint(0, rows, (n, fan), device=DEV, dtype=torch.int32)
    cnt = torch.zeros(n, device=DEV, dtype=torch.int32) if cz else torch.full((n,), fan, device=DEV, dtype=torch.int32)
    n_dev = torch.tensor([n], device=DEV, dtype=torch.int64)
    Wt = (torch.randn(256, 2 * Fp, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(256, device=DEV)
    out = torch.empty(n, 256, device=DEV, dtype=torch.bfloat16)
    A = torch.empty(n, 2 * Fp, device=DEV, dtype=torch.bfloat16)
    trace = torch.zeros(512, dtype=torch.int64, device=DEV)
    for _ in range(3):
        sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True, p_drop=pd, seed=1, out=out, save_A=True, A_out=A, trace=trace)
    torch.cuda.synchronize()
    t = trace.cpu().tolist()
    t0 = t[0]
    us = lambda c: (c - t0) / 1900.0 if c else -1
    KC = 2 * Fp // 64
    print(f"== n={n} Fp={Fp} cnt0={cz} drop={pd}: kernel end {us(t[1]):.1f} us (CTA 0)")
    for it in range(8):
        if not t[144 + it]:
            break
        print(f" tile {it}: top {us(t[144+it]):6.1f} | prod " + " ".join(f"{us(t[16+it*16+k]):6.1f}" for k in range(KC)) +
              " | mma " + " ".join(f"{us(t[160+it*16+k]):6.1f}" for k in range(KC)) +
              f" | epi wait {us(t[300+it*4]):6.1f} got {us(t[301+it*4]):6.1f} done {us(t[302+it*4]):6.1f}")
