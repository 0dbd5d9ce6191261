#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/i_$name.log 2>&1; echo "rc=$? $name"; tail -n 4 gpurun_out/i_$name.log | cut -c1-400; }
run tests python -m pytest tests/test_ops_gpu.py tests/test_nn_parts.py tests/test_model_zoo.py -q --timeout 300 -p no:cacheprovider -x
DGLB200_SDDMM_U=2 run kern_u2 python benchmarks/bench_kernels.py --impl ours --out gpurun_out/kernels_ours_u2.json
DGLB200_SDDMM_U=4 run kern_u4 python benchmarks/bench_kernels.py --impl ours --out gpurun_out/kernels_ours_u4.json --graphs reddit
python benchmarks/bench_kernels.py --compare gpurun_out/kernels_ours_u2.json gpurun_out/kernels_ref.json --table gpurun_out/kernels_compare_u2.md | tail -n 2
python benchmarks/bench_kernels.py --compare gpurun_out/kernels_ours_u4.json gpurun_out/kernels_ref.json --table gpurun_out/kernels_compare_u4.md | tail -n 2
