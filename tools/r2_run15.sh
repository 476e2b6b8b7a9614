#!/bin/bash
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
for v in kvcp; do
  VRDX_LIB=build/ab/libvrdx_$v.so timeout 600 python tools/shape_sweep.py --log2n 24 25 28 --algos 2 1 --shapes 0 2 --kinds kv > $O/sweep_$v.txt 2>&1
done
timeout 600 python tools/shape_sweep.py --log2n 24 25 28 --algos 2 1 --shapes 0 2 --kinds kv > $O/sweep_kvsts.txt 2>&1
grep -H "2^2" $O/sweep_kv*.txt | sed 's/gpurun_out.r2o.sweep_//'
{
echo "# compute-sanitizer {memcheck,racecheck,synccheck} python tools/sanitize_probe.py onesweep rts onesweep_256x16 rts_512x16 -- round-2 kernels"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_probe.py onesweep rts onesweep_256x16 rts_512x16 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | head -20
done
} > $O/compute_sanitizer.txt 2>&1
cat $O/compute_sanitizer.txt
