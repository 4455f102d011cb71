#!/bin/bash
# One GPU call that produces every measured artefact under profiles/ for a
# round (run under gpurun; outputs go to gpurun_out/, copy what is to be judged
# into profiles/).  Usage: bash tools/profile_round.sh <tag>
tag=${1:-r02}
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
python tools/ncu_traffic.py $out/traffic_$tag.csv --launches-per-step 32 > $out/ntt_traffic_$tag.json 2>> $out/bench_$tag.err
# the four transform kernels at full size (no slicing), full metric set
VKHEL_SLICE_MIB=0 ncu --set full --clock-control none --import-source on -k regex:ntt_ -s 8 -c 4 \
    -f -o $out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python tools/ncu_summary.py $out/prof_$tag.ncu-rep > $out/ncu_full_$tag.txt 2>&1
python tools/sweep.py > $out/sweep_$tag.jsonl 2>/dev/null
python tools/elem_bench.py 2>/dev/null | tail -1 > $out/elem_$tag.json
python tools/pcie_probe.py 2>/dev/null | tail -1 > $out/pcie_$tag.json
python tools/small_shard.py 2>/dev/null > $out/small_shard_$tag.jsonl
python tools/kernel_ab.py vkhel_b200/lib/libvkhel.so > $out/kernel_ab_$tag.txt 2>&1
for a in "10 4096" "12 1024" "14 256" "16 64"; do
  build/bin/api_loop $a | tail -1; VKHEL_NO_DEFER=1 build/bin/api_loop $a | tail -1
done > $out/api_loop_$tag.jsonl
for a in "8 2048" "10 2048" "12 1024" "14 256" "16 64"; do
  build/bin/api_product $a | tail -1
  VKHEL_NO_FUSED_PRODUCT=1 build/bin/api_product $a | tail -1
  VKHEL_NO_DEFER=1 build/bin/api_product $a | tail -1
done > $out/api_product_$tag.jsonl
for a in "16 64 8" "14 256 8" "12 1024 8"; do
  build/bin/api_e2e $a | tail -1; VKHEL_NO_READAHEAD=1 build/bin/api_e2e $a | tail -1
done > $out/api_e2e_$tag.jsonl
ls -la $out | tail -25
