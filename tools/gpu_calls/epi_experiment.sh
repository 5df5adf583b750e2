#!/bin/bash
for m in 0 1 2 4 5 7; do
  echo "=== STYLER_TC_DBGMODE=$m (bit0 no stores, bit1 no tmem ld, bit2 no tma store)"
  STYLER_TC_DBGMODE=$m timeout 100 python tools/phase_timing.py 2>&1 | grep -E "^fc_plain|^qkv_plain|^ffn1" | cut -c1-60,170-260
done
