#!/bin/bash
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
( timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -6 $O/pytest_gpu.txt )
for v in kvearly; do
  VRDX_LIB=build/ab/libvrdx_$v.so timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 2 3 --kinds kv > $O/sweep_$v.txt 2>&1
done
timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 2 3 --kinds kv > $O/sweep_kvlate.txt 2>&1
grep -H "2^2[58]" $O/sweep_kv*.txt | sed 's/gpurun_out.r2j.sweep_//'
timeout 600 python tools/shape_sweep.py --log2n 18 19 20 21 22 --algos 1 --shapes 0 --reps 9 > $O/small.txt 2>&1; grep "2^" $O/small.txt
# experiments library: the x_ flavours of the parity suite
VRDX_LIB=build/ab/libvrdx_x.so timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "x_ and (distributions or small_n or indirect_count or graph)" > $O/pytest_experiments.txt 2>&1; tail -4 $O/pytest_experiments.txt
VRDX_LIB=build/ab/libvrdx_x.so timeout 600 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --experiment 7 > $O/sweep_x7.txt 2>&1
VRDX_LIB=build/ab/libvrdx_x.so timeout 600 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --experiment 8 > $O/sweep_x8.txt 2>&1
VRDX_LIB=build/ab/libvrdx_x.so timeout 600 python tools/shape_sweep.py --log2n 28 --algos 2 1 --shapes 0 --experiment 6 > $O/sweep_x6.txt 2>&1
grep -H "2^28" $O/sweep_x*.txt | sed 's/gpurun_out.r2j.sweep_//'
