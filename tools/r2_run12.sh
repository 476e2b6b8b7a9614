#!/bin/bash
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
for v in head cur head cur; do
  if [ $v = cur ]; then L=vulkan_radix_sort_b200/lib/libvrdx_b200.so; else L=build/ab/libvrdx_$v.so; fi
  VRDX_LIB=$L timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 2 --kinds keys >> $O/ab_$v.txt 2>&1
  VRDX_LIB=$L timeout 600 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --kinds kv >> $O/ab_$v.txt 2>&1
done
grep -H "2^2[58]" $O/ab_*.txt | sed 's/gpurun_out.r2l.ab_//'
bash tools/ncu_all_kernels.sh $O/kernels
python tools/dist_kernels_bench.py 29 2>&1 | grep -E "class_count|partition"
