#!/bin/bash
out=gpurun_out; mkdir -p $out
python -c "
import vkhel_b200 as vk, json
c=vk.Context(0)
c.probe_int_peaks()
print(json.dumps(c.probe_int_peaks()))
" 2>&1 | tee $out/probe_r2c.txt
python tools/kernel_ab.py vkhel_b200/lib/libvkhel.so build/variants/libvkhel_c7.so build/variants/libvkhel_z1x.so 2>&1 | tee $out/kernel_ab_r2c.txt
python bench.py --steps 200 --warmup 5 > $out/bench_r2c.json 2> $out/bench_r2c.err; tail -3 $out/bench_r2c.err; cut -c1-1500 $out/bench_r2c.json
VKHEL_SLICE_MIB=0 ncu --set full --clock-control none -k regex:ntt_ -s 8 -c 4 -f -o $out/prof_r2c python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python tools/ncu_summary.py $out/prof_r2c.ncu-rep > $out/ncu_full_r2c.txt 2>&1
grep -E "^==|duration|fmaheavy|alu_cycles|issue_active|warps_active|inst_executed.sum|eligible|stalled" $out/ncu_full_r2c.txt | cut -c1-140
