"""Isolated timing of the fused tcgen05 SAGE layer kernel (variations used to find what bounds it)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgl_b200.ops.fused import sage_layer_tc

DEV = "cuda"


def timeit(fn, iters=20):
    if "--once" in sys.argv:  # under ncu: just launch, the profiler reports gpu__time_duration per launch
        fn(); fn(); torch.cuda.synchronize()
        return 1.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    torch.manual_seed(0)
    for (rows, n, Fp, fan) in [(2_400_000, 8704, 256, 10), (20_000, 8704, 256, 10), (2_400_000, 59520, 128, 10),
                               (20_000, 59520, 128, 10), (2_400_000, 148 * 128 * 4, 128, 10)]:
        X = (torch.randn(rows, Fp, device=DEV) * 0.5).bfloat16()
        self_ids = torch.randint(0, rows, (n,), device=DEV, dtype=torch.int32)
        nbr = torch.randint(0, rows, (n, fan), device=DEV, dtype=torch.int32)
        cnt = torch.full((n,), fan, device=DEV, dtype=torch.int32)
        n_dev = torch.tensor([n], device=DEV, dtype=torch.int64)
        Wt = (torch.randn(256, 2 * Fp, device=DEV) * 0.05).bfloat16()
        bias = torch.randn(256, device=DEV)
        out = torch.empty(n, 256, device=DEV, dtype=torch.bfloat16)
        A = torch.empty(n, 2 * Fp, device=DEV, dtype=torch.bfloat16)
        for p_drop, save in itertools.product([0.0, 0.5], [False, True]):
            t = timeit(lambda: sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True, p_drop=p_drop, seed=1,
                                             out=out, save_A=save, A_out=A if save else None))
            gb = n * (fan + 1) * Fp * 2 / 1e9
            print(f"rows={rows:8d} n={n:6d} Fp={Fp} fan={fan} drop={p_drop} saveA={int(save)}: {t:7.1f} us  "
                  f"gather {gb / t * 1e6:7.1f} GB/s  gemm {2 * n * 2 * Fp * 256 / t / 1e6:7.1f} TFLOP/s")
        # zero-fanout: isolates pipeline + MMA + epilogue from the gather
        cnt0 = torch.zeros_like(cnt)
        t = timeit(lambda: sage_layer_tc(X, self_ids, nbr, cnt0, n_dev, Wt, bias, relu=True, out=out))
        print(f"   cnt=0 variant: {t:7.1f} us")


if __name__ == "__main__":
    main()
