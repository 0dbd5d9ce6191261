"""Grouped GEMM (segment_mm) on tcgen05 vs the reference's formulation (host loop of one cuBLAS GEMM per relation,
src/array/cuda/gather_mm.cu:201-285, with the seglen D2H copy of nn/pytorch/linear.py:205-207).
CUDA-event timing, 5 warm-up + 20 timed iterations; output one JSON line per shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import dgl_b200 as dgl


def timed(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = "cuda"
    res = []
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6462.4)
    for (N, R, K, M, skew) in [(2_000_000, 64, 128, 128, 0.0), (2_000_000, 64, 128, 128, 1.2), (1_000_000, 8, 256, 256, 0.0),
                               (4_000_000, 349, 64, 64, 1.0), (500_000, 16, 512, 512, 0.0)]:
        g = torch.Generator().manual_seed(0)
        w = torch.ones(R) if skew == 0 else 1.0 / torch.arange(1, R + 1).float() ** skew
        seglen = torch.bincount(torch.multinomial(w, N, replacement=True, generator=g), minlength=R).to(dev)
        a = torch.randn(N, K, device=dev).bfloat16()
        b = (torch.randn(R, K, M, device=dev) * 0.1).bfloat16()
        a.requires_grad_(), b.requires_grad_()
        go = torch.randn(N, M, device=dev).bfloat16()

        def ours_fwd():
            with torch.no_grad():
                return dgl.ops.segment_mm(a, b, seglen)

        def ours_fwdbwd():
            a.grad = b.grad = None
            dgl.ops.segment_mm(a, b, seglen).backward(go)

        def ref_loop_fwd():  # the reference: seglen to the host, one GEMM per relation
            with torch.no_grad():
                out = torch.empty(N, M, device=dev, dtype=a.dtype)
                off = 0
                for r, n in enumerate(seglen.tolist()):
                    if n:
                        torch.matmul(a[off:off + n], b[r], out=out[off:off + n])
                    off += n
                return out

        def ref_loop_fwdbwd():
            with torch.no_grad():
                out = torch.empty(N, M, device=dev, dtype=a.dtype)
                da = torch.empty_like(a)
                db = torch.empty_like(b)
                off = 0
                for r, n in enumerate(seglen.tolist()):
                    if n:
                        torch.matmul(a[off:off + n], b[r], out=out[off:off + n])
                        torch.matmul(go[off:off + n], b[r].t(), out=da[off:off + n])
                        torch.matmul(a[off:off + n].t(), go[off:off + n], out=db[r])
                    else:
                        db[r].zero_()
                    off += n
                return out

        err = (ours_fwd().float() - ref_loop_fwd().float()).abs().max().item()
        t1, t2 = timed(ours_fwd), timed(ref_loop_fwd)
        t3, t4 = timed(ours_fwdbwd), timed(ref_loop_fwdbwd)
        flops = 2.0 * N * K * M
        byts = (N * K + N * M) * 2 + R * K * M * 2
        rec = dict(bench="segment_mm", N=N, R=R, K=K, M=M, zipf_skew=skew, dtype="bf16",
                   ms_tc_fwd=t1, ms_cublas_loop_fwd=t2, speedup_fwd=t2 / t1,
                   ms_tc_fwd_bwd=t3, ms_cublas_loop_fwd_bwd=t4, speedup_fwd_bwd=t4 / t3,
                   tflops_fwd=flops / t1 / 1e9, gbps_fwd=byts / t1 / 1e6, frac_hbm_fwd=byts / t1 / 1e6 / hbm,
                   max_abs_diff_vs_cublas=err)
        print(json.dumps(rec), flush=True)
        res.append(rec)
        del a, b, go
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
