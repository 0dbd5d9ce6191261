#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CASES=u_dot_v:128:float32,u_dot_v:128:bfloat16,u_mul_e_sum:602:float32,copy_u_sum:602:float32,u_mul_e_sum:128:float32
timeout 600 ncu --set full --clock-control none -k 'regex:spmm|sddmm' -f -o /tmp/kern_ours2 \
   python benchmarks/ncu_kernels.py --impl ours --cases $CASES > gpurun_out/s_ncu_ours.log 2>&1
echo "rc=$? ncu"; tail -n 2 gpurun_out/s_ncu_ours.log
ncu -i /tmp/kern_ours2.ncu-rep --page raw --csv > gpurun_out/kern_ours2_raw.csv 2>/dev/null
timeout 300 python benchmarks/bench_kernels.py --impl ours --out gpurun_out/s_kernels_ours.json > gpurun_out/s_kern.log 2>&1; echo "rc=$? kern"
ls -la gpurun_out/kern_ours2_raw.csv
