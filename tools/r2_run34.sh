#!/bin/bash
mkdir -p gpurun_out/r2ah
O=gpurun_out/r2ah
VRDX_LIB=build/ab/libvrdx_up5.so timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 --kinds kv keys > $O/sweep_up5.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 --kinds kv keys > $O/sweep_base.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2ah.sweep_//'
