#!/bin/bash
# Developer sweep: time every compiled pass-kernel variant (keys-only and key-value) at 2^26 and 2^28.
NV=${1:-8}
for v in $(seq 0 $((NV-1))); do
  echo "=== variant $v"
  VRDX_KEYS_VARIANT=$v VRDX_KV_VARIANT=$v timeout 120 python tools/quick_bench.py --log2n 26 28 --reps 3 2>&1 | grep -E "tile|GKeys|sorted|Error|error"
done
