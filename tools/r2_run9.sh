#!/bin/bash
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
for g in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2953$g bench.py --gpus $g --steps 5 --warmup 3 > $O/bench_dist$g.json 2> $O/bench_dist$g.err; echo "rc=$?"; tail -c 300 $O/bench_dist$g.err
python -c "
import json; d=json.loads(open('$O/bench_dist$g.json').read().strip().splitlines()[-1])
print('$g GPUs value',d['value'],'ms',d['ms_per_step'],d['stages_ms_max_over_ranks'],'exch GB/s',d['exchange_gbs_per_gpu'], 'verified', d['config']['verified'])
print({k:(v['verified'],v['received_min'],v['received_max']) for k,v in d['config']['adversarial_parity'].items()}); print(d['single_gpu_sort_of_one_ranks_input'])
"
done
