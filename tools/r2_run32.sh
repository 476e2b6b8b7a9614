#!/bin/bash
mkdir -p gpurun_out/r2af
O=gpurun_out/r2af
VRDX_LIB=build/ab/libvrdx_nolazy.so timeout 300 python tools/shape_sweep.py --log2n 25 28 29 --algos 2 --shapes 0 --kinds keys > $O/sweep_nolazy.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 25 28 29 --algos 2 --shapes 0 --kinds keys > $O/sweep_lazy.txt 2>&1
VRDX_LIB=build/ab/libvrdx_nolazy.so timeout 300 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --kinds keys > $O/sweep_nolazy_b.txt 2>&1
timeout 300 python tools/shape_sweep.py --log2n 28 --algos 2 --shapes 0 --kinds keys > $O/sweep_lazy_b.txt 2>&1
grep -H "2^2\|WRONG" $O/sweep_*.txt | sed 's/gpurun_out.r2af.sweep_//'
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sort_ex_gpu.py -x -q -m gpu -k "block_free or distributions or full_size or adversarial or tail_tile" > $O/pytest.txt 2>&1; tail -3 $O/pytest.txt
