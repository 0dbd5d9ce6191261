#!/bin/bash
# first GPU pass of round 2: new tcgen05 kernels (unit tests), engine, sampling fix, both bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 300 "$@" > gpurun_out/a_$name.log 2>&1; echo "rc=$? $name"; tail -n 6 gpurun_out/a_$name.log; }
run fwd32 python -m pytest tests/test_sage_bwd_tc_gpu.py -q -k "layer_tc_fp32" --timeout 120 -p no:cacheprovider
run fwd16 python -m pytest tests/test_sage_tc_gpu.py -q --timeout 120 -p no:cacheprovider
run dw python -m pytest tests/test_sage_bwd_tc_gpu.py -q -k "bwd_dw" --timeout 120 -p no:cacheprovider
run da python -m pytest tests/test_sage_bwd_tc_gpu.py -q -k "bwd_da" --timeout 120 -p no:cacheprovider
run autograd python -m pytest tests/test_sage_bwd_tc_gpu.py -q -k "autograd" --timeout 120 -p no:cacheprovider
run engine python -m pytest tests/test_engine_gpu.py -q --timeout 200 -p no:cacheprovider
run sampling python -m pytest tests/test_sampling_gpu.py -q --timeout 120 -p no:cacheprovider
run bench_ours python bench.py --steps 20 --warmup 5
run bench_ref python bench.py --impl reference --steps 20 --warmup 5
