#!/bin/bash
# 2-GPU pass: NVLS multicast path, data-parallel tcgen05 engine, N=2 bench with NVLS on / off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/d_topo.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 500 "$@" > gpurun_out/d_$name.log 2>&1; echo "rc=$? $name"; tail -n 6 gpurun_out/d_$name.log; }
run nvls python -m pytest tests/test_multigpu.py -q -s -k "nvls or trainer_tc" --timeout 300 -p no:cacheprovider
run mgpu python -m pytest tests/test_multigpu.py -q -k "not nvls and not trainer_tc" --timeout 300 -p no:cacheprovider
for mode in 1 0; do
  echo "=== bench2 NVLS=$mode"
  DGLB200_NVLS=$mode timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
     bench.py --gpus 2 --steps 20 --warmup 5 --no-epoch > gpurun_out/d_bench2_nvls$mode.log 2>&1
  echo "rc=$?"; tail -n 2 gpurun_out/d_bench2_nvls$mode.log | cut -c1-1500
done
echo "=== ref2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
     bench.py --impl reference --gpus 2 --steps 20 --warmup 5 --no-epoch > gpurun_out/d_ref2.log 2>&1
echo "rc=$?"; tail -n 1 gpurun_out/d_ref2.log | cut -c1-1200
