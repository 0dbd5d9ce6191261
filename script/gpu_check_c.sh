#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 400 "$@" > gpurun_out/c_$name.log 2>&1; echo "rc=$? $name"; tail -n 6 gpurun_out/c_$name.log; }
run bwd python -m pytest tests/test_sage_bwd_tc_gpu.py -q --timeout 120 -p no:cacheprovider
run engine python -m pytest tests/test_engine_gpu.py -q --timeout 300 -p no:cacheprovider
run bench python bench.py --steps 20 --warmup 5
for d in bf16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv \
    --log-file gpurun_out/c_launch_$d.csv python bench.py --steps 3 --warmup 3 --no-graph --dtype $d --no-epoch > gpurun_out/c_ncu_$d.log 2>&1
  echo "--- $d"; python tools/ncu_summarize.py gpurun_out/c_launch_$d.csv 2>&1 | head -14
done
