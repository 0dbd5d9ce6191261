#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CASES=u_dot_v:128:float32,u_dot_v:128:bfloat16,u_mul_e_sum:602:float32,copy_u_sum:602:float32,u_mul_e_sum:128:float32
for impl in ours reference; do
  timeout 900 ncu --set full --clock-control none -k 'regex:spmm|sddmm|SpMM|SDDMM' -f -o /tmp/kern_$impl \
     python benchmarks/ncu_kernels.py --impl $impl --cases $CASES > gpurun_out/h_ncu_$impl.log 2>&1
  echo "rc=$? ncu $impl"; tail -n 3 gpurun_out/h_ncu_$impl.log
  ncu -i /tmp/kern_$impl.ncu-rep --page raw --csv > gpurun_out/kern_${impl}_raw.csv 2>/dev/null
done
ls -la gpurun_out/kern_*
