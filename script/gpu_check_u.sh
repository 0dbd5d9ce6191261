#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 400 "$@" > gpurun_out/u_$name.log 2>&1; echo "rc=$? $name"; tail -n 3 gpurun_out/u_$name.log | cut -c1-500; }
run tests python -m pytest tests/test_graphbolt_gpu.py -q --timeout 300 -p no:cacheprovider -k "graphed or capacity"
rm -f gpurun_out/u_pipe.jsonl
P="python benchmarks/bench_uva_pipeline.py --scale 0.05 --warmup 10 --out gpurun_out/u_pipe.jsonl"
run gat_graphed $P --impl ours --resident device --model gat --graphed --steps 40
run gat_graphed_amp $P --impl ours --resident device --model gat --graphed --amp --steps 40
