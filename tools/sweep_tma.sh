#!/bin/bash
# Developer sweep of the persistent TMA-staged variants (keys-only and key-value) at 2^28.
NV=${1:-8}
for v in $(seq 0 $((NV-1))); do
  echo "=== tma variant $v"
  VRDX_KEYS_TMA_VARIANT=$v VRDX_KV_TMA_VARIANT=$v timeout 120 python tools/quick_bench.py --log2n 28 --reps 3 2>&1 | grep -E "tile|GKeys|sorted|Error|error"
done
echo "=== direct (tile_load=1)"
VRDX_TILE_LOAD=1 timeout 120 python tools/quick_bench.py --log2n 28 --reps 3 2>&1 | grep -E "tile|GKeys|sorted|Error|error"
