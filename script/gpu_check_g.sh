#!/bin/bash
# bulk-copy SpMM + host-path edits: tests, kernel bench (both arms), pipeline re-measure
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/g_$name.log 2>&1; echo "rc=$? $name"; tail -n 6 gpurun_out/g_$name.log | cut -c1-600; }
run tests python -m pytest tests/test_ops_gpu.py tests/test_sage_bwd_tc_gpu.py tests/test_sampling_gpu.py tests/test_graphbolt_gpu.py tests/test_layer_plans.py -q --timeout 300 -p no:cacheprovider
run kern_ours python benchmarks/bench_kernels.py --impl ours --out gpurun_out/kernels_ours.json
DGLB200_SPMM_BULK=0 run kern_ours_nobulk python benchmarks/bench_kernels.py --impl ours --out gpurun_out/kernels_ours_nobulk.json
run kern_ref python benchmarks/bench_kernels.py --impl reference --out gpurun_out/kernels_ref.json
python benchmarks/bench_kernels.py --compare gpurun_out/kernels_ours.json gpurun_out/kernels_ref.json --table gpurun_out/kernels_compare.md | tail -n 3
python benchmarks/bench_kernels.py --compare gpurun_out/kernels_ours.json gpurun_out/kernels_ours_nobulk.json --table gpurun_out/kernels_bulk_vs_simt.md | tail -n 3
rm -f gpurun_out/g_pipe.jsonl
run pipe_dev_ours python benchmarks/bench_uva_pipeline.py --impl ours --resident device --scale 0.05 --steps 60 --warmup 10 --workers 0 --out gpurun_out/g_pipe.jsonl
run pipe_uva_ours python benchmarks/bench_uva_pipeline.py --impl ours --resident host --scale 0.05 --steps 60 --warmup 10 --workers 1 --out gpurun_out/g_pipe.jsonl
run hostprof python benchmarks/profile_pipeline.py
