#!/bin/bash
# round-2 GPU call 1: parity of the new tile kernel + A/B of its ranking variants and tile shapes + one ncu capture
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt )
for v in r1p1 r0p1 r2p1 r1p0 r0p0 r1p1nf r0p1nf; do
  VRDX_LIB=build/ab/libvrdx_$v.so timeout 600 python tools/shape_sweep.py --log2n 28 > $O/sweep_$v.txt 2>&1
  grep -c WRONG $O/sweep_$v.txt
done
VRDX_LIB=build/ab/libvrdx_x.so timeout 600 python tools/shape_sweep.py --log2n 28 --shapes 0 --experiment 6 > $O/sweep_x_round1.txt 2>&1
VRDX_LIB=build/ab/libvrdx_r1p1.so timeout 600 python tools/shape_sweep.py --log2n 20 22 24 25 26 --shapes 0 1 > $O/sweep_r1p1_small.txt 2>&1
# ncu: scatter pass of the product library (keys, stable pass = 2nd launch of the kernel)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:PassKernel -s 1 -c 1 -o $O/prof_scatter_keys python tools/ncu_one.py 28 keys 1 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:PassKernel -s 1 -c 1 -o $O/prof_scatter_kv python tools/ncu_one.py 28 kv 1 > $O/ncu2.log 2>&1
grep -h "2^28" $O/sweep_*.txt | sort -k9 -n -r | head -60
