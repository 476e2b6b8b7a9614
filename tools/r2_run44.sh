#!/bin/bash
mkdir -p gpurun_out/r2ar
O=gpurun_out/r2ar
VRDX_LIB=build/ab/libvrdx_grow.so timeout 300 python tools/shape_sweep.py --log2n 20 22 24 26 28 --algos 1 --shapes 0 --kinds keys kv > $O/sweep_grow.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 20 22 24 26 28 --algos 1 --shapes 0 --kinds keys kv > $O/sweep_base.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2ar.sweep_//'
