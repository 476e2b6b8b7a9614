#!/bin/bash
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > $O/smi.txt
timeout 300 python tools/shape_sweep.py --log2n 26 27 28 --algos 2 1 --shapes 0 --kinds kv keys > $O/sweep_a.txt 2>&1
timeout 600 python tools/sweep_n.py "1<<26" "1<<27" "1<<28" > $O/sweep_n.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 26 27 28 --algos 1 2 --shapes 0 --kinds kv > $O/sweep_b.txt 2>&1
cat $O/smi.txt; grep -h "2^2" $O/sweep_a.txt; cat $O/sweep_n.txt; grep -h "2^2" $O/sweep_b.txt
