#!/bin/bash
mkdir -p gpurun_out/r2ac
O=gpurun_out/r2ac
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
{
echo "# compute-sanitizer {memcheck,racecheck,synccheck} VRDX_TWO_RUNS=1 python tools/sanitize_probe.py onesweep rts onesweep_256x16 rts_512x16"
echo "# (final round-2 kernels: block-free one-run / two-run tiles, fused spine kernel with ticketed segments, 128-bit upsweep)"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  VRDX_TWO_RUNS=1 timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_probe.py onesweep rts onesweep_256x16 rts_512x16 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error|Assert" | head -20
done
} > $O/compute_sanitizer.txt 2>&1
cat $O/compute_sanitizer.txt
timeout 300 python tools/shape_sweep.py --log2n 25 28 --algos 2 --shapes 0 --kinds keys kv > $O/sweep.txt 2>&1; grep -h "2^2" $O/sweep.txt
