#!/bin/bash
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "block_free or distributions or all_flavours or full_size or adversarial" > $O/pytest_parity.txt 2>&1; tail -3 $O/pytest_parity.txt
for v in two0 nobf; do
  VRDX_LIB=build/ab/libvrdx_$v.so timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 --kinds keys kv > $O/sweep_$v.txt 2>&1
done
timeout 300 python tools/shape_sweep.py --log2n 25 26 27 28 29 --algos 2 --shapes 0 --kinds keys kv > $O/sweep_new.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2u.sweep_//'
