#!/bin/bash
mkdir -p gpurun_out/r2al
timeout 400 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 1 2 3 --kinds keys > gpurun_out/r2al/sweep.txt 2>&1; grep -h "2^2\|WRONG" gpurun_out/r2al/sweep.txt
