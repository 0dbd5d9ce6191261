#!/bin/bash
# 2-GPU box: partition tests, then the partitioned bench at N=1 (ours) and N=2 (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/l_$name.log 2>&1; echo "rc=$? $name"; tail -n 3 gpurun_out/l_$name.log | cut -c1-1100; }
run tests python -m pytest tests/test_multigpu.py -q --timeout 600 -p no:cacheprovider -x -k "partitioned or cooperative"
run part_ours_1 python bench.py --mode partitioned --gpus 1 --steps 5 --warmup 2
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
run part_ours_2 $L bench.py --mode partitioned --gpus 2 --steps 5 --warmup 2
DGLB200_HALO_MODE=fused run part_ours_2_fused $L bench.py --mode partitioned --gpus 2 --steps 5 --warmup 2
run part_ref_2 $L bench.py --mode partitioned --impl reference --gpus 2 --steps 5 --warmup 2
grep -h '"metric"' gpurun_out/l_part_*.log > gpurun_out/partitioned_l.jsonl
