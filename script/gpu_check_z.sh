#!/bin/bash
# Re-entry validation: full GPU suite, compute-sanitizer memcheck over a kernel subset, short both-arm bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; t=$2; shift 2; echo "=== $name"; timeout $t "$@" > gpurun_out/z_$name.log 2>&1; echo "rc=$? $name"; tail -n 2 gpurun_out/z_$name.log | cut -c1-400; }
run tests 240 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider -x
run sanitizer 200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -q -p no:cacheprovider -x \
  "tests/test_visit_topk_gpu.py::test_topk_matches_bincount_oracle[40-5-30-idt0]" \
  "tests/test_visit_topk_gpu.py::test_topk_matches_bincount_oracle[3000-16-100000-idt1]" \
  tests/test_graph_kernels_gpu.py -k "topk or coo_to_csr or coalesce or spgemm or random_walk or labor0 or partition_ops"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/z_sanitizer.log | tail -n 5
run bench_ref 200 python bench.py --impl reference --steps 30 --warmup 5
run bench_ours 120 python bench.py --steps 50 --warmup 5
