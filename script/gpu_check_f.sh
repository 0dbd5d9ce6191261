#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/f_$name.log 2>&1; echo "rc=$? $name"; tail -n 5 gpurun_out/f_$name.log | cut -c1-600; }
run samp python -m pytest tests/test_sampling_gpu.py tests/test_graphbolt_gpu.py tests/test_graph_kernels_gpu.py tests/test_layer_plans.py -q --timeout 300 -p no:cacheprovider
rm -f gpurun_out/f_pipe.jsonl
for impl in ours reference; do
  run pipe_dev_$impl python benchmarks/bench_uva_pipeline.py --impl $impl --resident device --scale 0.05 --steps 60 --warmup 10 --workers 0 --out gpurun_out/f_pipe.jsonl
done
run pipe_dev_ours_amp python benchmarks/bench_uva_pipeline.py --impl ours --resident device --scale 0.05 --steps 60 --warmup 10 --workers 0 --amp --out gpurun_out/f_pipe.jsonl
run pipe_uva_ours python benchmarks/bench_uva_pipeline.py --impl ours --resident host --scale 0.05 --steps 60 --warmup 10 --workers 1 --out gpurun_out/f_pipe.jsonl
run pipe_uva_ref python benchmarks/bench_uva_pipeline.py --impl reference --resident host --scale 0.05 --steps 60 --warmup 10 --out gpurun_out/f_pipe.jsonl
