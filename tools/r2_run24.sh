#!/bin/bash
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
timeout 600 python tools/sweep_n.py "1<<24" "3<<23" "1<<25" "3<<24" "1<<26" "3<<25" "1<<27" "3<<26" "1<<28" > $O/sweep_n.txt 2>&1; cat $O/sweep_n.txt
