#!/bin/bash
# block-free tiles (TileBlockFree) A/B + parity
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu > $O/pytest_parity.txt 2>&1; tail -3 $O/pytest_parity.txt
VRDX_LIB=build/ab/libvrdx_nobf.so timeout 300 python tools/shape_sweep.py --log2n 25 26 28 --algos 2 --shapes 0 1 --kinds keys > $O/sweep_nobf.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 25 26 28 --algos 2 --shapes 0 1 --kinds keys kv > $O/sweep_bf.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2p.sweep_//'
