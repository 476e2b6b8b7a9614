#!/bin/bash
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
( timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -8 $O/pytest_gpu.txt )
for rt in 4 8 16 32; do
  VRDX_RANGE_TILES=$rt timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 1 2 --kinds keys > $O/sweep_rt${rt}_keys.txt 2>&1
  VRDX_RANGE_TILES=$rt timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 2 --kinds kv > $O/sweep_rt${rt}_kv.txt 2>&1
done
VRDX_RTS_PERSISTENT=0 timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 1 2 --kinds keys > $O/sweep_pertile_keys.txt 2>&1
VRDX_RTS_PERSISTENT=0 timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 2 --kinds kv > $O/sweep_pertile_kv.txt 2>&1
grep -H "2^2[58]" $O/sweep_*.txt | sed 's/gpurun_out.r2c.sweep_//' | sort -k7,7 -k8,8 -k10,10n
