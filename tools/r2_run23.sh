#!/bin/bash
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
VRDX_LIB=build/ab/libvrdx_ldg128.so timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 --kinds keys > $O/sweep_ldg128.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 --kinds keys > $O/sweep_ldg32.txt 2>&1
VRDX_LIB=build/ab/libvrdx_ldg128.so timeout 300 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --kinds keys > $O/sweep_ldg128_b.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --kinds keys > $O/sweep_ldg32_b.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2w.sweep_//'
timeout 600 python tools/sweep_n.py > $O/sweep_n.txt 2>&1; cat $O/sweep_n.txt
