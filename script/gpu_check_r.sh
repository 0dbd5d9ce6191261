#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/r_$name.log 2>&1; echo "rc=$? $name"; tail -n 2 gpurun_out/r_$name.log | cut -c1-420; }
run tests python -m pytest tests/test_ops_gpu.py tests/test_layer_plans.py tests/test_graphbolt_gpu.py -q --timeout 300 -p no:cacheprovider
rm -f gpurun_out/r_pipe.jsonl
P="python benchmarks/bench_uva_pipeline.py --scale 0.05 --warmup 10 --out gpurun_out/r_pipe.jsonl"
run gat_fused $P --impl ours --resident device --model gat --steps 40 --workers 0
DGLB200_GAT_FUSED=0 run gat_unfused $P --impl ours --resident device --model gat --steps 40 --workers 0
run uva_cache $P --impl ours --resident host --workers 1 --cache-mb 1024 --steps 60
run kern python benchmarks/bench_kernels.py --impl ours --out gpurun_out/r_kernels_ours.json --graphs ogbn-arxiv,reddit
SCALE=0.05 MODEL=gat timeout 200 python benchmarks/profile_pipeline.py > gpurun_out/r_prof_gat.log 2>&1
