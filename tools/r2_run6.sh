#!/bin/bash
# 2-GPU call: P2P store microbenchmark, distributed bench baseline (+ adversarial parity, NVLink counters), two-device test
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
nvidia-smi topo -m > $O/topo.txt 2>&1
nvidia-smi nvlink -gt d -i 0 > $O/nvlink_sample.txt 2>&1; head -8 $O/nvlink_sample.txt
./tools/microbench_p2p > $O/p2p.txt 2>&1; cat $O/p2p.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "two_devices or two_host_threads" > $O/pytest_two.txt 2>&1; tail -3 $O/pytest_two.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_dist2.json 2> $O/bench_dist2.err; tail -c 400 $O/bench_dist2.err
python -c "
import json; d=json.loads(open('$O/bench_dist2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],d['stages_ms_max_over_ranks'],'exch GB/s',d['exchange_gbs_per_gpu'])
print(d['config']['adversarial_parity']); print(d['nvlink_counters']); print(d['single_gpu_sort_of_one_ranks_input'])
"
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q > $O/pytest_dist.txt 2>&1; tail -3 $O/pytest_dist.txt
