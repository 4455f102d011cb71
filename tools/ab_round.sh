#!/bin/bash
# A/B of library variants on the GPU box (run under gpurun): parity first, then
# bench.py per variant (twice, alternating), then the sweep with the default build.
# Usage: bash tools/ab_round.sh <tag> <variant.so>...
tag=${1:-ab}; shift
out=gpurun_out; mkdir -p $out
( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_tables.py \
    tests/test_gpu_deferred.py tests/test_gpu_edge_cases.py -x -q -m gpu ) \
    > $out/pytest_$tag.log 2>&1
tail -5 $out/pytest_$tag.log
for rep in 1 2; do
  bash tools/bench_variants.sh vkhel_b200/lib/libvkhel.so "$@"
done | tee $out/variants_$tag.txt
python tools/sweep.py > $out/sweep_$tag.jsonl 2>/dev/null
grep -E "sweep|elem|polymul|2\^14" $out/sweep_$tag.jsonl | cut -c1-330
