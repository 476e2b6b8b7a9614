#!/bin/bash
# N GPUs: distributed tests (N=2 only) + bench.py --gpus N
N=${1:-2}
O=gpurun_out/r2ad
mkdir -p $O
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q > $O/pytest_dist.txt 2>&1; tail -3 $O/pytest_dist.txt; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_dist$N.json 2> $O/bench_dist$N.err; tail -c 300 $O/bench_dist$N.err
python - <<PY
import json
d=json.load(open("$O/bench_dist$N.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus")})
print(json.dumps(d.get("extra",{}))[:1500])
PY
