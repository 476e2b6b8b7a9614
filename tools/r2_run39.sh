#!/bin/bash
mkdir -p gpurun_out/r2am
O=gpurun_out/r2am
for v in kvr4 kvr8 kvr12 base; do
  VRDX_LIB=build/ab/libvrdx_$v.so timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 1 --shapes 0 --kinds kv > $O/sweep_$v.txt 2>&1
done
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2am.sweep_//'
