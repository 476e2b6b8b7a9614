#!/bin/bash
# block-free tiles (one / two runs) + fused spine: parity and A/B
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "block_free or distributions or all_flavours or full_size or adversarial or graph" > $O/pytest_parity.txt 2>&1; tail -3 $O/pytest_parity.txt
VRDX_LIB=build/ab/libvrdx_spine2.so timeout 300 python tools/shape_sweep.py --log2n 25 26 28 --algos 2 --shapes 0 --kinds keys kv > $O/sweep_spine2.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 25 26 27 28 29 --algos 2 --shapes 0 1 --kinds keys kv > $O/sweep_new.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2r.sweep_//'
