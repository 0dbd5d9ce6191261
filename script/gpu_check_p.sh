#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/p_$name.log 2>&1; echo "rc=$? $name"; tail -n 3 gpurun_out/p_$name.log | cut -c1-700; }
run tests python -m pytest tests/test_graphbolt_gpu.py tests/test_sage_bwd_tc_gpu.py tests/test_graph_kernels_gpu.py -q --timeout 300 -p no:cacheprovider
rm -f gpurun_out/p_pipe.jsonl
run pipe_table python benchmarks/bench_uva_pipeline.py --impl ours --resident device --graphed --table-input --scale 0.05 --steps 60 --warmup 10 --out gpurun_out/p_pipe.jsonl
run pipe_table_amp python benchmarks/bench_uva_pipeline.py --impl ours --resident device --graphed --table-input --amp --scale 0.05 --steps 60 --warmup 10 --out gpurun_out/p_pipe.jsonl
run pipe_gat_ours python benchmarks/bench_uva_pipeline.py --impl ours --resident device --model gat --scale 0.05 --steps 40 --warmup 10 --workers 0 --out gpurun_out/p_pipe.jsonl
run pipe_sage_ours python benchmarks/bench_uva_pipeline.py --impl ours --resident device --scale 0.05 --steps 60 --warmup 10 --workers 0 --out gpurun_out/p_pipe.jsonl
SCALE=0.05 GRAPHED=1 TABLE=1 timeout 200 python benchmarks/profile_pipeline.py > gpurun_out/p_prof_table.log 2>&1
