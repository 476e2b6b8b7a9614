#!/bin/bash
# 2-GPU call: new splitter stage, partition probe under ncu (NVLink bytes), dist tests
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
for st in sampled exact; do
VRDX_DIST_SPLITTERS=$st timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_dist2_$st.json 2> $O/bench_dist2_$st.err; tail -c 300 $O/bench_dist2_$st.err
python -c "
import json; d=json.loads(open('$O/bench_dist2_$st.json').read().strip().splitlines()[-1])
print('$st value',d['value'],'ms',d['ms_per_step'],d['stages_ms_max_over_ranks'],'exch GB/s',d['exchange_gbs_per_gpu'], 'verified', d['config']['verified'])
print({k:(v['verified'],v['received_min'],v['received_max']) for k,v in d['config']['adversarial_parity'].items()})
"
done
./tools/p2p_partition_probe 2 | tee $O/probe2.txt
./tools/p2p_partition_probe 8 | tee $O/probe8.txt
timeout 300 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:DistPartition -c 2 ./tools/p2p_partition_probe 8 > $O/ncu_nvlink.txt 2>&1; grep -E "nvl|gpu__time|dram__|DistPartition" $O/ncu_nvlink.txt | head -20
timeout 300 ncu --query-metrics 2>/dev/null | grep -i "^nvl" | head -20 > $O/nvl_metrics.txt; head -5 $O/nvl_metrics.txt
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q > $O/pytest_dist.txt 2>&1; tail -3 $O/pytest_dist.txt
python tools/dist_kernels_bench.py 29 2>&1 | tail -12
