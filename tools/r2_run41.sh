#!/bin/bash
mkdir -p gpurun_out/r2ao
VRDX_LIB=build/ab/libvrdx_shapes.so timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 1 --shapes 0 4 5 6 7 --kinds kv --pairs-only-shape > gpurun_out/r2ao/sweep.txt 2>&1; grep -h "2^2\|WRONG\|Error\|error" gpurun_out/r2ao/sweep.txt
