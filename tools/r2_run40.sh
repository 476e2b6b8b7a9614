#!/bin/bash
mkdir -p gpurun_out/r2an
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sort_ex_gpu.py -x -q -m gpu -k "distributions or adversarial or low_entropy or all_flavours or idempotent" > gpurun_out/r2an/pytest.txt 2>&1; tail -3 gpurun_out/r2an/pytest.txt
timeout 300 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --kinds kv keys --dist all_zero > gpurun_out/r2an/sweep_allzero.txt 2>&1; grep -h "2^2\|WRONG" gpurun_out/r2an/sweep_allzero.txt
