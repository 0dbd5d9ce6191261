#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x -k "not multigpu" > gpurun_out/e_full.log 2>&1
echo "rc=$?"; tail -n 25 gpurun_out/e_full.log | cut -c1-300
