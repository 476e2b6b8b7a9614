#!/bin/bash
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
for v in prod nopf; do
  VRDX_LIB=build/ab/libvrdx_$v.so VRDX_RANGE_TILES=8 timeout 600 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 1 2 --kinds keys > $O/sweep_$v.txt 2>&1
  VRDX_LIB=build/ab/libvrdx_$v.so VRDX_RANGE_TILES=8 timeout 600 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --kinds kv >> $O/sweep_$v.txt 2>&1
done
grep -H "2^28" $O/sweep_*.txt | sed 's/gpurun_out.r2d.sweep_//'
VRDX_RANGE_TILES=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:RangePassKernel -s 1 -c 1 -o $O/prof_range_keys python tools/ncu_one.py 28 keys 1 > $O/ncu1.log 2>&1
VRDX_RANGE_TILES=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:RangePassKernel -s 1 -c 1 -o $O/prof_range_kv python tools/ncu_one.py 28 kv 1 > $O/ncu2.log 2>&1
ls -la $O
