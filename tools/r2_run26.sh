#!/bin/bash
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 300 python tools/shape_sweep.py --log2n 25 26 27 28 29 --algos 2 --shapes 0 --kinds kv keys > $O/sweep_a.txt 2>&1
grep -h "2^2\|WRONG" $O/sweep_a.txt
timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "block_free or distributions or all_flavours or full_size or adversarial or four_byte or offsets or indirect" > $O/pytest_parity.txt 2>&1; tail -3 $O/pytest_parity.txt
