#!/bin/bash
# Re-validation after the visit-counter rewrite and the pinned-staging change: tests, racecheck, two benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/zy_pinsage.jsonl gpurun_out/zy_fetch.jsonl
run() { name=$1; t=$2; shift 2; echo "=== $name"; timeout $t "$@" > gpurun_out/zy_$name.log 2>&1; echo "rc=$? $name"; tail -n 3 gpurun_out/zy_$name.log | cut -c1-260; }
run tests 200 python -m pytest tests/test_visit_topk_gpu.py tests/test_graphbolt_async.py tests/test_graphbolt_gpu.py -q --timeout 150 -p no:cacheprovider -x
run racecheck 150 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest -q -p no:cacheprovider -x \
  "tests/test_visit_topk_gpu.py::test_topk_matches_bincount_oracle[700-10-200-idt1]" \
  "tests/test_visit_topk_gpu.py::test_topk_matches_bincount_oracle[40-5-30-idt0]"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/zy_racecheck.log | tail -n 4
run pin_ours 150 python benchmarks/bench_pinsage_sampler.py --impl ours --out gpurun_out/zy_pinsage.jsonl
run fetch 240 python benchmarks/bench_async_fetch.py --out gpurun_out/zy_fetch.jsonl
