#!/bin/bash
# Tuning sweep of the tcgen05 conv kernel (stages via smem budget, N tile) on the config-3 shapes.
for kb in 110 160 210; do
  for bn in 0 128; do
    echo "=== STYLER_TC_SMEM_KB=$kb STYLER_TC_BN=$bn"
    STYLER_TC_SMEM_KB=$kb STYLER_TC_BN=$bn timeout 120 python tools/prof_kernels.py --iters 3 2>&1 | grep -E "ffn1|ffn2_ln|qkv|fc_ln|postnet1|attention"
  done
done
