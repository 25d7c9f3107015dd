#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>  ->  gpurun_out/variants/lib_<name>.so (A/B experiments)
name=$1; shift
mkdir -p /tmp/variants/$name gpurun_out/variants
cd arpeggio_b200/csrc
for f in arp_api arp_pairs arp_planes arp_sifts arp_json arp_rings; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC "$@" -c $f.cu -o /tmp/variants/$name/$f.o || exit 1
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$name.so /tmp/variants/$name/*.o -cudart static
