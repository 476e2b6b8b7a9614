#!/bin/bash
# A/B: two-run tiles on/off x table alignment 256/16
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
for v in two0 align16 two0align16; do
  VRDX_LIB=build/ab/libvrdx_$v.so timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 --kinds keys kv > $O/sweep_$v.txt 2>&1
done
timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 --kinds keys kv > $O/sweep_new.txt 2>&1
VRDX_LIB=build/ab/libvrdx_two0align16.so timeout 300 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --kinds keys > $O/sweep_two0align16_again.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2s.sweep_//'
