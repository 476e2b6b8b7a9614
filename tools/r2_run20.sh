#!/bin/bash
O=gpurun_out/r2t
mkdir -p $O
M="--set full --clock-control none --import-source on"
run() {  # name, kernel regex, launches to skip, env, command...
  local name=$1 regex=$2 skip=$3 envs=$4; shift 4
  env $envs timeout 600 ncu $M -k regex:$regex -s $skip -c 1 -o $O/$name "$@" > /dev/null 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep 100 > $O/$name.txt 2>&1
  ncu -i $O/$name.ncu-rep --page source --csv > $O/$name.src.csv 2>/dev/null
  python tools/ncu_source_ops.py $O/$name.src.csv ${KEYS:-2**28} > $O/$name.ops.txt 2>/dev/null
  rm -f $O/$name.ncu-rep
  grep -E "kernel:|time_duration|dram__bytes_(read|write).sum \[" $O/$name.txt | head -4
}
KEYS=2**28 run k_pass0_two1 PassKernel 0 "X=1" python tools/ncu_one.py 28 keys 1
KEYS=2**28 run k_pass2_two1 PassKernel 2 "X=1" python tools/ncu_one.py 28 keys 1
KEYS=2**28 run k_upsweep2_two1 UpsweepKernel 2 "X=1" python tools/ncu_one.py 28 keys 1
