#!/bin/bash
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
( timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -6 $O/pytest_gpu.txt )
timeout 600 python tools/shape_sweep.py --log2n 18 19 20 21 22 24 --algos 1 --shapes 0 2 --reps 9 > $O/small.txt 2>&1; grep "2^" $O/small.txt
timeout 600 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 2 --kinds keys > $O/big.txt 2>&1; grep "2^" $O/big.txt
timeout 300 ./bench_cpp/bench b200 --sizes 2^18,2^19,2^20,2^21,2^22 --seed 1 --runs 10 -o $O/b200_small.csv | tail -12
timeout 300 ./bench_cpp/bench cuda --sizes 2^18,2^19,2^20,2^21,2^22 --seed 1 --runs 10 --no-verify -o $O/cuda_small.csv | tail -12
