#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/n_$name.log 2>&1; echo "rc=$? $name"; tail -n 4 gpurun_out/n_$name.log | cut -c1-700; }
run tests python -m pytest tests/test_ops_gpu.py tests/test_sampling_gpu.py tests/test_graphbolt_gpu.py -q --timeout 300 -p no:cacheprovider -k "dot or labor or graphed or staged"
rm -f gpurun_out/n_pipe.jsonl
run pipe_graphed python benchmarks/bench_uva_pipeline.py --impl ours --resident device --graphed --scale 0.05 --steps 60 --warmup 10 --out gpurun_out/n_pipe.jsonl
run pipe_graphed_amp python benchmarks/bench_uva_pipeline.py --impl ours --resident device --graphed --amp --scale 0.05 --steps 60 --warmup 10 --out gpurun_out/n_pipe.jsonl
run pipe_gat_ours python benchmarks/bench_uva_pipeline.py --impl ours --resident device --model gat --scale 0.05 --steps 40 --warmup 10 --workers 0 --out gpurun_out/n_pipe.jsonl
run pipe_gat_ref python benchmarks/bench_uva_pipeline.py --impl reference --resident device --model gat --scale 0.05 --steps 40 --warmup 10 --out gpurun_out/n_pipe.jsonl
