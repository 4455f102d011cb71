#!/bin/bash
# build_variant.sh NAME "-DFOO=1 -DBAR=2"  ->  build/variants/libvkhel_NAME.so
# (kernel tuning experiments; load with VKHEL_LIB_PATH=...)
set -e
cd "$(dirname "$0")/.."
name=$1; defs=$2
out=build/variants; mkdir -p $out/obj_$name
for f in device vector kernels_elem kernels_ntt kernels_ntt_cluster kernels_ntt_tma kernels_ntt_small tables_device probe hostcopy; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -Iinclude -Iinclude/vkhel -Ivkhel_b200/csrc $defs \
    -c vkhel_b200/csrc/$f.cu -o $out/obj_$name/$f.o &
done
wait
for f in numbers ntt_tables; do
  /usr/bin/gcc -O2 -fPIC -Iinclude -Iinclude/vkhel -c vkhel_b200/csrc/$f.c -o $out/obj_$name/$f.o
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libvkhel_$name.so \
  $out/obj_$name/*.o -Xlinker --version-script=vkhel.syms -Xlinker -z -Xlinker nodelete
echo built $out/libvkhel_$name.so
