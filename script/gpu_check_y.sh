#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 400 "$@" > gpurun_out/y_$name.log 2>&1; echo "rc=$? $name"; tail -n 1 gpurun_out/y_$name.log | cut -c1-330; }
run tests python -m pytest tests/test_layer_plans.py tests/test_graphbolt_gpu.py -q --timeout 300 -p no:cacheprovider
rm -f gpurun_out/y_pipe.jsonl gpurun_out/y_configs.jsonl
P="python benchmarks/bench_uva_pipeline.py --scale 0.05 --warmup 10 --steps 40 --resident device --model gat --out gpurun_out/y_pipe.jsonl"
run gat_ours $P --impl ours --workers 0
run gat_ours_amp $P --impl ours --amp --workers 0
run gat_graphed_amp $P --impl ours --amp --graphed
run gat_graphed $P --impl ours --graphed
run gat_reddit python benchmarks/bench_configs.py --config gat_reddit --impl ours --out gpurun_out/y_configs.jsonl
