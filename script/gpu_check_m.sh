#!/bin/bash
# N-GPU box: partitioned bench at N (ours; reference too when REF=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/m_$name.log 2>&1; echo "rc=$? $name"; tail -n 1 gpurun_out/m_$name.log | cut -c1-1300; }
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
run part_ours_$N $L bench.py --mode partitioned --gpus $N --steps 10 --warmup 3
if [ "$REF" = "1" ]; then run part_ref_$N $L bench.py --mode partitioned --impl reference --gpus $N --steps 5 --warmup 2; fi
if [ "$TESTS" = "1" ]; then run tests python -m pytest tests/test_multigpu.py -q --timeout 600 -p no:cacheprovider; fi
grep -h '"metric"' gpurun_out/m_part_*_$N.log > gpurun_out/partitioned_m$N.jsonl
