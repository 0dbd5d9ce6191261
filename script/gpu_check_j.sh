#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}
run() { name=$1; shift; echo "=== $name"; timeout 1500 "$@" > gpurun_out/j_$name.log 2>&1; echo "rc=$? $name"; tail -n 4 gpurun_out/j_$name.log | cut -c1-900; }
if [ "$N" = "1" ]; then
  run tests python -m pytest tests/test_ops_gpu.py -q --timeout 300 -p no:cacheprovider -x -k "dot or sddmm or spmm"
  run part_ours_1 python bench.py --mode partitioned --gpus 1 --steps 5 --warmup 2
  run part_ref_1 python bench.py --mode partitioned --impl reference --gpus 1 --steps 5 --warmup 2
else
  L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
  run part_ours_$N $L bench.py --mode partitioned --gpus $N --steps 5 --warmup 2
  run part_ref_$N $L bench.py --mode partitioned --impl reference --gpus $N --steps 5 --warmup 2
fi
grep -h '"metric"' gpurun_out/j_part_*_$N.log >> gpurun_out/partitioned.jsonl
