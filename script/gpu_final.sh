#!/bin/bash
# Final 1-GPU evidence run: full GPU suite, headline bench (both arms), kernel bench (both arms), configs 3/4/5, pipelines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/final
O=gpurun_out/final
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > $O/$name.log 2>&1; echo "rc=$? $name"; tail -n 2 $O/$name.log | cut -c1-600; }
run tests python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_ref python bench.py --impl reference --gpus 1 --steps 100 --warmup 10
run bench_ours python bench.py --gpus 1 --steps 100 --warmup 10
run kern_ours python benchmarks/bench_kernels.py --impl ours --out $O/kernels_ours.json
run kern_ref python benchmarks/bench_kernels.py --impl reference --out $O/kernels_ref.json
python benchmarks/bench_kernels.py --compare $O/kernels_ours.json $O/kernels_ref.json --table $O/kernels_compare.md | tail -n 2
rm -f $O/configs.jsonl $O/pipe.jsonl
for impl in ours reference; do
  run gat_$impl python benchmarks/bench_configs.py --config gat_reddit --impl $impl --out $O/configs.jsonl
  run rgcn_$impl python benchmarks/bench_configs.py --config rgcn_mag --impl $impl --out $O/configs.jsonl
done
P="python benchmarks/bench_uva_pipeline.py --scale 0.05 --steps 60 --warmup 10 --out $O/pipe.jsonl"
run pipe_dev_ref $P --impl reference --resident device
run pipe_dev_ours $P --impl ours --resident device --workers 0
run pipe_dev_graphed $P --impl ours --resident device --graphed
run pipe_dev_table $P --impl ours --resident device --graphed --table-input
run pipe_dev_table_amp $P --impl ours --resident device --graphed --table-input --amp
run pipe_gat_ref $P --impl reference --resident device --model gat --steps 40
run pipe_gat_ours $P --impl ours --resident device --model gat --steps 40 --workers 0
run pipe_uva_ref $P --impl reference --resident host
run pipe_uva_ours $P --impl ours --resident host --workers 1
run pipe_uva_ours_cache $P --impl ours --resident host --workers 1 --cache-mb 1024
grep -h '"metric"' $O/bench_ref.log $O/bench_ours.log > $O/bench.jsonl
