#!/bin/bash
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "block_free or distributions or all_flavours or full_size or adversarial" > $O/pytest_parity.txt 2>&1; tail -3 $O/pytest_parity.txt
timeout 300 python tools/shape_sweep.py --log2n 25 26 27 28 29 --algos 2 --shapes 0 --kinds keys > $O/sweep_new.txt 2>&1
VRDX_TWO_RUNS=0 timeout 300 python tools/shape_sweep.py --log2n 25 26 27 28 29 --algos 2 --shapes 0 --kinds keys > $O/sweep_two_never.txt 2>&1
VRDX_TWO_RUNS=1 timeout 300 python tools/shape_sweep.py --log2n 25 26 27 28 29 --algos 2 --shapes 0 --kinds keys > $O/sweep_two_always.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2v.sweep_//'
