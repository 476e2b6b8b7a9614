#!/bin/bash
# One ncu --set full capture per kernel on the path, summarised on the box (developer tool; run under gpurun).
M="--set full --clock-control none --import-source on"
cap() {  # name, kernel regex, skip, env, command...
  local name=$1 regex=$2 skip=$3; shift 3
  env "$@" > /dev/null 2>&1
}
run() {
  local name=$1 regex=$2 skip=$3 envs=$4; shift 4
  env $envs ncu $M -k regex:$regex -s $skip -c 1 -o gpurun_out/$name "$@" > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep 100 > gpurun_out/$name.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
  grep -E "kernel:|time_duration|dram__bytes_(read|write).sum \[" gpurun_out/$name.txt | head -4
}
run k_upsweep UpsweepKernel 1 "X=1" python tools/ncu_one.py 28 keys 1
run k_spine_reduce SpineReduce 1 "X=1" python tools/ncu_one.py 28 keys 1
run k_spine_apply SpineApply 1 "X=1" python tools/ncu_one.py 28 keys 1
run k_scatter_kv OnesweepKernel 1 "X=1" python tools/ncu_one.py 28 kv 1
run k_hist_private HistogramKernelPrivate 0 "VRDX_ALGORITHM=1" python tools/ncu_one.py 24 keys 1
run k_onesweep_lookback OnesweepKernel 1 "VRDX_ALGORITHM=1" python tools/ncu_one.py 24 keys 1
run k_dist_partition DistPartition 0 "X=1" python tools/dist_kernels_bench.py 27
run k_dist_hist DistPrefixHistogram 4 "X=1" python tools/dist_kernels_bench.py 27
