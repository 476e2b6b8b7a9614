#!/bin/bash
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 900 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 5 8 9 10 11 12 --kinds keys > $O/ipt_keys.txt 2>&1
timeout 900 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 6 8 9 10 11 12 --kinds kv > $O/ipt_kv.txt 2>&1
timeout 900 python tools/shape_sweep.py --log2n 21 22 23 24 --algos 1 --shapes 0 5 9 10 11 > $O/ipt_one.txt 2>&1
grep -h "2^" $O/ipt_*.txt
