#!/usr/bin/env python
"""g-SpMM / g-SDDMM kernel benchmark, both arms on the same box and the same graphs.

Metric conventions of the reference's ASV kernels (/root/reference/benchmarks/benchmarks/kernel/bench_gspmm_copy_u.py:10-39,
bench_gspmm_u_mul_e_sum.py, bench_gsddmm_u_dot_v.py): 3 warm-up + 10 timed iterations, CUDA events, "GFLOPS" = E*F/t for
copy_u and 2*E*F/t for u_mul_e / u_dot_v. Between timed iterations a 256 MB buffer is written to evict L2.

  python benchmarks/bench_kernels.py --impl ours       [--scale 1.0] [--out gpurun_out/kernels_ours.json]
  python benchmarks/bench_kernels.py --impl reference  [--scale 1.0] [--out gpurun_out/kernels_ref.json]
  python benchmarks/bench_kernels.py --compare gpurun_out/kernels_ours.json gpurun_out/kernels_ref.json

Graphs: RMAT with the node / edge counts of ogbn-arxiv, reddit (1/4 of its edges), ogbn-products, generated with plain
torch from a fixed seed so both arms see the identical edge list. F in {32, 100, 128, 256, 602}, fp32 and bf16.
The bytes model for the roofline column is the ALGORITHMIC traffic (E*4 index bytes + E*F*s gathered rows + N*F*s
output); the DRAM-side numbers come from the ncu captures in profiles/.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {"ogbn-arxiv": (169343, 1166243, False), "reddit": (232965, 114615892 // 4, False),
          "ogbn-products": (2449029, 61859140, True)}
FEATS = (32, 100, 128, 256, 602)


def rmat_edges(n, e, sym, device, seed=0, a=0.57, b=0.19, c=0.19):
    scale = max(1, (n - 1).bit_length())
    gen = torch.Generator(device=device).manual_seed(seed)
    src = torch.zeros(e, dtype=torch.int64, device=device)
    dst = torch.zeros(e, dtype=torch.int64, device=device)
    for _ in range(scale):
        r = torch.rand(e, device=device, generator=gen)
        src = (src << 1) | (r >= a + b).long()
        dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c)).long()
    src, dst = (src * 0x9E3779B1 + 0x7F4A7C15) % (1 << scale) % n, (dst * 0x9E3779B1 + 0x7F4A7C15) % (1 << scale) % n
    if sym:
        src, dst = torch.cat([src, dst]), torch.cat([dst, src])
    return src.int(), dst.int()


def timeit(fn, flush, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def run(args):
    if args.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        os.environ.setdefault("DGLBACKEND", "pytorch")
        import dgl
    else:
        sys.path.insert(0, ROOT)
        import dgl_b200 as dgl
    ops = dgl.ops
    peaks = {"hbm_gbs": 6462.4}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    results = []
    for name in args.graphs.split(","):
        n, e, sym = SHAPES[name]
        n, e = int(n * args.scale), int(e * args.scale)
        u, v = rmat_edges(n, e, sym, dev)
        g = dgl.graph((u, v), num_nodes=n, idtype=torch.int32, device=dev)
        del u, v
        g.create_formats_()
        E = g.num_edges()
        for F in FEATS:
            for dtype in (torch.float32, torch.bfloat16):
                x = torch.randn(n, F, device=dev).to(dtype)
                w = torch.rand(E, 1, device=dev).to(dtype)
                s = x.element_size()
                base = dict(graph=name, N=n, E=E, F=F, dtype=str(dtype).replace("torch.", ""), impl=args.impl)
                cases = [("gspmm_copy_u_sum", lambda: ops.copy_u_sum(g, x), E * 4 + E * F * s + n * F * s, E * F),
                         ("gspmm_u_mul_e_sum", lambda: ops.u_mul_e_sum(g, x, w), E * 4 + E * F * s + n * F * s + E * s, 2 * E * F)]
                if E * F * s < 24e9:
                    cases.append(("gsddmm_u_dot_v", lambda: ops.u_dot_v(g, x, x), E * 8 + 2 * E * F * s + E * s, 2 * E * F))
                for kname, fn, nbytes, flops in cases:
                    try:
                        t = timeit(fn, flush)
                    except Exception as ex:  # noqa: BLE001 - e.g. an arm without bf16 support for an op
                        results.append(dict(base, kernel=kname, error=f"{type(ex).__name__}: {ex}"[:120]))
                        continue
                    results.append(dict(base, kernel=kname, ms=t, GBps=nbytes / t / 1e6,
                                        frac_of_measured_hbm=nbytes / t / 1e6 / peaks["hbm_gbs"],
                                        GFLOPS_ref_convention=flops / t / 1e6))
                del x, w
            print(json.dumps(results[-1]), flush=True)
        del g
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump({"peaks": peaks, "impl": args.impl, "results": results}, open(args.out, "w"), indent=1)


def compare(ours_path, ref_path, out=None):
    ours = {(r["kernel"], r["graph"], r["F"], r["dtype"]): r for r in json.load(open(ours_path))["results"] if "ms" in r}
    ref = {(r["kernel"], r["graph"], r["F"], r["dtype"]): r for r in json.load(open(ref_path))["results"] if "ms" in r}
    lines = ["| kernel | graph | F | dtype | ours ms | reference ms | speed-up | ours GB/s (algorithmic) |", "|---|---|---|---|---|---|---|---|"]
    worst = None
    for k in sorted(ours):
        o, r = ours[k], ref.get(k)
        if r is None:
            continue
        sp = r["ms"] / o["ms"]
        worst = sp if worst is None else min(worst, sp)
        lines.append(f"| {k[0]} | {k[1]} | {k[2]} | {k[3]} | {o['ms']:.3f} | {r['ms']:.3f} | {sp:.2f}x | {o['GBps']:.0f} |")
    lines.append(f"\nsmallest speed-up over the reference kernels: {worst:.2f}x" if worst is not None else "\nno common rows")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--graphs", default="ogbn-arxiv,reddit,ogbn-products")
    ap.add_argument("--compare", nargs=2, default=None, metavar=("OURS_JSON", "REF_JSON"))
    ap.add_argument("--table", default=None, help="with --compare: also write the markdown table here")
    args = ap.parse_args()
    if args.compare:
        compare(args.compare[0], args.compare[1], args.table)
        return
    if args.out is None:
        args.out = f"gpurun_out/kernels_{args.impl}.json"
    run(args)


if __name__ == "__main__":
    main()
