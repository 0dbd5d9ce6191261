#!/bin/bash
# ncu --set full captures of the engine step kernels (one per kernel instance), both dtypes, + sampler A/B + host profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in bf16 fp32; do
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
     -k regex:"sage_layer_tc|sage_bwd_d|frontier_compact|sample_ell|allreduce_adam|ce_loss2" -c 17 -f -o gpurun_out/prof_engine_$d \
     python bench.py --steps 1 --warmup 3 --no-graph --dtype $d --no-epoch > gpurun_out/p_ncu_$d.log 2>&1
  echo "ncu $d rc=$?"
  ncu -i gpurun_out/prof_engine_$d.ncu-rep --page raw --csv > gpurun_out/prof_engine_${d}_raw.csv 2>/dev/null
done
echo "=== sampler A/B on the pinned (UVA) graph"
for m in 1 0; do
  DGLB200_FUSED_SAMPLER=$m timeout 600 python benchmarks/bench_uva_pipeline.py --impl ours --resident host --scale 0.05 --steps 40 --warmup 8 --workers 1 2>/dev/null | tail -1 | cut -c1-400
done
echo "=== host profile of the generic pipeline"
SCALE=0.02 timeout 600 python benchmarks/profile_pipeline.py > gpurun_out/p_hostprofile.log 2>&1; echo rc=$?
head -60 gpurun_out/p_hostprofile.log | cut -c1-200
