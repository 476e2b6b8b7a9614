#!/bin/bash
mkdir -p gpurun_out/r2ai
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "crossovers" > gpurun_out/r2ai/pytest.txt 2>&1; tail -3 gpurun_out/r2ai/pytest.txt
