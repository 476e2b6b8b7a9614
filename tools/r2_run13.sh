#!/bin/bash
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 2 5 6 7 --kinds keys > $O/ipt_keys.txt 2>&1
timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 5 6 --kinds kv > $O/ipt_kv.txt 2>&1
timeout 600 python tools/shape_sweep.py --log2n 22 24 --algos 1 --shapes 0 2 5 7 > $O/ipt_one.txt 2>&1
grep -h "2^" $O/ipt_*.txt
timeout 600 python tools/shape_sweep.py --log2n 18 20 --algos 1 --shapes 0 --reps 9 --kinds keys > $O/small.txt 2>&1; grep "2^" $O/small.txt
