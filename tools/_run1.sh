set -x
nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" 
for t in 1 2 3 4 6 8; do echo "== copy threads $t"; VKHEL_COPY_THREADS=$t build/bin/hostcopy_bench | tail -2; done
for t in 1 2 3 4 6 8; do echo "== api_e2e threads $t"; VKHEL_COPY_THREADS=$t build/bin/api_e2e 16 64 10; done
for t in 1 4; do VKHEL_COPY_THREADS=$t build/bin/api_e2e 14 256 8;  VKHEL_COPY_THREADS=$t build/bin/api_e2e 17 32 8; done
python tools/kernel_ab.py vkhel_b200/lib/libvkhel.so build/variants/libvkhel_fwdlazy.so
