#!/bin/bash
# Developer sweep of tile shapes for the reduce-then-scan composition.
for v in 0 8 9 10 11 12 13 14 15; do
  echo "=== rts variant $v"
  VRDX_ALGORITHM=2 VRDX_KEYS_VARIANT=$v VRDX_KV_VARIANT=$v timeout 120 python tools/quick_bench.py --log2n 28 --reps 3 2>&1 | grep -E "tile|GKeys|sorted|Error|error"
done
