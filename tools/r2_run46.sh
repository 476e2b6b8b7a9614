#!/bin/bash
mkdir -p gpurun_out/r2as
timeout 400 python tools/fuzz_parity.py 120 3 21 > gpurun_out/r2as/fuzz3.txt 2>&1; tail -3 gpurun_out/r2as/fuzz3.txt
VRDX_TWO_RUNS=1 timeout 400 python tools/fuzz_parity.py 80 4 23 > gpurun_out/r2as/fuzz4.txt 2>&1; tail -3 gpurun_out/r2as/fuzz4.txt
