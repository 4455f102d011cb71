#!/bin/bash
# One GPU call that produces every measured artefact under profiles/ for a
# round (run under gpurun; outputs go to gpurun_out/, copy what is to be judged
# into profiles/).  Usage: bash tools/profile_round.sh <tag>
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
# every launch of a short run with its device time (recipe: cold cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
# DRAM traffic with the caches left alone (one pass per kernel)
ncu --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    -c 400 --csv --log-file $out/traffic_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
# the four transform kernels at full size (no slicing), full metric set
VKHEL_SLICE_MIB=0 ncu --set full --clock-control none -k regex:ntt_ -s 8 -c 4 \
    -f -o $out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
# the fused polynomial-product kernel
ncu --set full --clock-control none -k regex:polymul -c 1 \
    -f -o $out/prof_polymul_$tag python tools/sweep.py --log2-total 24 > /dev/null 2>&1
# text summaries made on the box; the polymul report itself is not kept
python tools/ncu_summary.py $out/prof_$tag.ncu-rep > $out/ncu_full_$tag.txt 2>&1
python tools/ncu_summary.py $out/prof_polymul_$tag.ncu-rep > $out/ncu_full_polymul_$tag.txt 2>&1
rm -f $out/prof_polymul_$tag.ncu-rep
python tools/sweep.py > $out/sweep_$tag.jsonl 2>/dev/null
python tools/elem_bench.py 2>/dev/null | tail -1 > $out/elem_$tag.json
# one element-wise kernel, full metric set (HBM-bound: DRAM throughput)
ncu --set full --clock-control none -k regex:elem_vec -s 6 -c 2 \
    -f -o $out/prof_elem_$tag python tools/elem_bench.py > /dev/null 2>&1
python tools/ncu_summary.py $out/prof_elem_$tag.ncu-rep > $out/ncu_full_elem_$tag.txt 2>&1
rm -f $out/prof_elem_$tag.ncu-rep
python tools/pcie_probe.py 2>/dev/null | tail -1 > $out/pcie_$tag.json
python tools/tables_bench.py 2>/dev/null | grep config > $out/tables_$tag.jsonl
for a in "10 4096" "12 1024" "14 256" "16 64"; do
  build/bin/api_loop $a | tail -1; VKHEL_NO_DEFER=1 build/bin/api_loop $a | tail -1
done > $out/api_loop_$tag.jsonl
for a in "8 2048" "10 2048" "12 1024" "14 256" "16 64"; do
  build/bin/api_product $a | tail -1
  VKHEL_NO_FUSED_PRODUCT=1 build/bin/api_product $a | tail -1
  VKHEL_NO_DEFER=1 build/bin/api_product $a | tail -1
done > $out/api_product_$tag.jsonl
build/bin/bfly_bench > $out/bfly_bench_$tag.txt 2>&1
build/bin/pipe_bench > $out/pipe_bench_$tag.txt 2>&1
build/bin/exchange_bench > $out/exchange_bench_$tag.txt 2>&1
ls -la $out | tail -20
