#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/zw_fetch.jsonl
run() { name=$1; t=$2; shift 2; echo "=== $name"; timeout $t "$@" > gpurun_out/zw_$name.log 2>&1; echo "rc=$? $name"; tail -n 2 gpurun_out/zw_$name.log | cut -c1-700; }
run fetch 200 python benchmarks/bench_async_fetch.py --out gpurun_out/zw_fetch.jsonl
run bench_ours 120 python bench.py --steps 100 --warmup 10
