#!/bin/bash
O=gpurun_out/r2ak
mkdir -p $O
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q > $O/pytest_dist.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_dist.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_dist2.json 2> $O/bench_dist2.err; echo "bench rc=$?"; tail -c 600 $O/bench_dist2.err; head -c 300 $O/bench_dist2.json
