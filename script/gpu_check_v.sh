#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 400 "$@" > gpurun_out/v_$name.log 2>&1; echo "rc=$? $name"; tail -n 1 gpurun_out/v_$name.log | cut -c1-330; }
run tests python -m pytest tests/test_layer_plans.py tests/test_graphbolt_gpu.py -q --timeout 300 -p no:cacheprovider
rm -f gpurun_out/v_pipe.jsonl
P="python benchmarks/bench_uva_pipeline.py --scale 0.05 --warmup 10 --steps 40 --resident device --model gat --out gpurun_out/v_pipe.jsonl"
run gat_ref_amp $P --impl reference --amp
run gat_ours_amp $P --impl ours --amp --workers 0
run gat_graphed_amp $P --impl ours --amp --graphed
run gat_graphed $P --impl ours --graphed
