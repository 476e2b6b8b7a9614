#!/bin/bash
N=${1:-2}
O=gpurun_out/r2ad
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_dist$N.json 2> $O/bench_dist$N.err; echo "rc=$?"; tail -c 1500 $O/bench_dist$N.err; head -c 2500 $O/bench_dist$N.json
