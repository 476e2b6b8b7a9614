#!/bin/bash
mkdir -p gpurun_out/r2as
timeout 400 python tools/fuzz_parity.py 150 1 > gpurun_out/r2as/fuzz1.txt 2>&1; tail -3 gpurun_out/r2as/fuzz1.txt
VRDX_TWO_RUNS=1 timeout 400 python tools/fuzz_parity.py 100 2 > gpurun_out/r2as/fuzz2.txt 2>&1; tail -3 gpurun_out/r2as/fuzz2.txt
