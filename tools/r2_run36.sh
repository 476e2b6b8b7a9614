#!/bin/bash
mkdir -p gpurun_out/r2aj
VRDX_LIB=build/ab/libvrdx_exp.so timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "small_n_sweep or distributions or all_flavours or block_free or indirect_count" > gpurun_out/r2aj/pytest_exp.txt 2>&1; tail -3 gpurun_out/r2aj/pytest_exp.txt
