#!/bin/bash
mkdir -p gpurun_out/r2aq
O=gpurun_out/r2aq
for v in lw2 lw8; do
  VRDX_LIB=build/ab/libvrdx_$v.so timeout 300 python tools/shape_sweep.py --log2n 22 24 26 --algos 1 --shapes 0 --kinds keys kv > $O/sweep_$v.txt 2>&1
done
timeout 300 python tools/shape_sweep.py --log2n 22 24 26 --algos 1 --shapes 0 --kinds keys kv > $O/sweep_lw4.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2aq.sweep_//'
