#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/zv_fetch.jsonl
timeout 200 python benchmarks/bench_async_fetch.py --out gpurun_out/zv_fetch.jsonl > gpurun_out/zv_fetch.log 2>&1; echo rc=$?
cut -c1-160 gpurun_out/zv_fetch.jsonl
