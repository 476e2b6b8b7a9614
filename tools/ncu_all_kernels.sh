#!/bin/bash
# One ncu --set full capture per kernel on the path, summarised on the box (developer tool; run under gpurun).
# usage: tools/ncu_all_kernels.sh <output dir under gpurun_out/>
O=${1:-gpurun_out/kernels}
mkdir -p $O
M="--set full --clock-control none --import-source on"
run() {  # name, kernel regex, launches to skip, env, command...
  local name=$1 regex=$2 skip=$3 envs=$4; shift 4
  env $envs timeout 600 ncu $M -k regex:$regex -s $skip -c 1 -o $O/$name "$@" > /dev/null 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep 100 > $O/$name.txt 2>&1
  ncu -i $O/$name.ncu-rep --page source --csv > $O/$name.src.csv 2>/dev/null
  python tools/ncu_source_ops.py $O/$name.src.csv ${KEYS:-2**28} > $O/$name.ops.txt 2>/dev/null
  rm -f $O/$name.ncu-rep $O/$name.src.csv
  grep -E "kernel:|time_duration|dram__bytes_(read|write).sum \[" $O/$name.txt | head -4
}
KEYS=2**28 run k_scatter_keys_pass0_blockfree PassKernel 0 "X=1" python tools/ncu_one.py 28 keys 1
KEYS=2**28 run k_scatter_keys_pass2_two_runs PassKernel 2 "X=1" python tools/ncu_one.py 28 keys 1
KEYS=2**28 run k_scatter_keys_pass3_stable PassKernel 3 "X=1" python tools/ncu_one.py 28 keys 1
KEYS=2**28 run k_scatter_kv PassKernel 1 "VRDX_ALGORITHM=2" python tools/ncu_one.py 28 kv 1
KEYS=2**28 run k_upsweep UpsweepKernel 1 "X=1" python tools/ncu_one.py 28 keys 1
KEYS=2**28 run k_upsweep_kv UpsweepKernel 1 "VRDX_ALGORITHM=2" python tools/ncu_one.py 28 kv 1
KEYS=2**28 run k_spine SpineKernel 1 "X=1" python tools/ncu_one.py 28 keys 1
KEYS=2**24 run k_hist_private HistogramKernelPrivate 0 "VRDX_ALGORITHM=1" python tools/ncu_one.py 24 keys 1
KEYS=2**24 run k_onesweep_lookback PassKernel 1 "VRDX_ALGORITHM=1" python tools/ncu_one.py 24 keys 1
[ -n "$SKIP_DIST" ] && exit 0
KEYS=2**29 run k_dist_class_count DistClassCount 0 "X=1" python tools/dist_kernels_bench.py 29
KEYS=2**29 run k_dist_partition DistPartition 1 "X=1" python tools/dist_kernels_bench.py 29
