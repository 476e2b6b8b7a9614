#!/bin/bash
# Developer sweep of tile shapes for the reduce-then-scan composition.
for v in ${@:-0 1 8 9 10 11}; do
  echo "=== rts variant $v"
  VRDX_ALGORITHM=2 VRDX_KEYS_RTS_VARIANT=$v VRDX_KV_RTS_VARIANT=$v timeout 120 python tools/quick_bench.py --log2n 25 28 --reps 3 2>&1 | grep -E "GKeys|sorted|Error|error"
done
