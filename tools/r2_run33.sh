#!/bin/bash
mkdir -p gpurun_out/r2ag
O=gpurun_out/r2ag
VRDX_TWO_RUNS=1 timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_sort_ex_gpu.py -x -q -m gpu > $O/pytest_two_always.txt 2>&1; tail -3 $O/pytest_two_always.txt
VRDX_ALGORITHM=2 timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_sort_ex_gpu.py -x -q -m gpu -k "not launch_count and not storage_reuse" > $O/pytest_force_rts.txt 2>&1; tail -3 $O/pytest_force_rts.txt
