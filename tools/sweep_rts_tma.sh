#!/bin/bash
for v in 0 1 2 3 4 5 6 7; do
  echo "=== rts+tma variant $v"
  VRDX_ALGORITHM=2 VRDX_TILE_LOAD=2 VRDX_KEYS_TMA_VARIANT=$v VRDX_KV_TMA_VARIANT=$v timeout 120 python tools/quick_bench.py --log2n 28 --reps 3 2>&1 | grep -E "tile|GKeys|sorted|Error|error"
done
