#!/bin/bash
# PinSAGE sampler (three arms), overlapped feature fetch, one ncu capture of the visit counter.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/zz_pinsage.jsonl gpurun_out/zz_fetch.jsonl
run() { name=$1; t=$2; shift 2; echo "=== $name"; timeout $t "$@" > gpurun_out/zz_$name.log 2>&1; echo "rc=$? $name"; tail -n 3 gpurun_out/zz_$name.log | cut -c1-300; }
run pin_ours 150 python benchmarks/bench_pinsage_sampler.py --impl ours --out gpurun_out/zz_pinsage.jsonl
run pin_sort 150 python benchmarks/bench_pinsage_sampler.py --impl ours-sort --out gpurun_out/zz_pinsage.jsonl
run pin_ref 200 python benchmarks/bench_pinsage_sampler.py --impl reference --out gpurun_out/zz_pinsage.jsonl
run fetch 240 python benchmarks/bench_async_fetch.py --out gpurun_out/zz_fetch.jsonl
run ncu 200 ncu --set full --clock-control none --import-source on -k regex:visit_topk -c 2 -o gpurun_out/zz_visit_topk \
  python benchmarks/bench_pinsage_sampler.py --impl ours --configs wide --steps 2 --warmup 1
ncu -i gpurun_out/zz_visit_topk.ncu-rep --page raw --csv > gpurun_out/zz_visit_topk_raw.csv 2>/dev/null
rm -f gpurun_out/zz_visit_topk.ncu-rep
