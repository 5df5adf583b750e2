#!/bin/bash
# round-2 GPU call B: new kernels first (pairs, bilstm), then the full suite, then A/B timings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "cta_pairs or bilstm or conv1d_tc_persistent" > $O/r2b_new.log 2>&1; echo "new rc=$?" >> $O/r2b_new.log
grep -E "passed|failed|FAILED|Error|rc=" $O/r2b_new.log | tail -15
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/pytest_gpu.log | tail -30
for m in 0 1; do
  STYLER_TC_2CTA=$m timeout 200 python tools/prof_kernels.py --only ffn1,postnet1,ffn2_ln,fc_ln,pred_conv_ln,bilstm_h80 > $O/r2b_prof_2cta$m.txt 2>&1
  echo "== 2CTA=$m"; cat $O/r2b_prof_2cta$m.txt
done
for m in 0 1; do
  STYLER_TC_2CTA=$m timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > $O/r2b_bench_2cta$m.json 2>$O/r2b_bench_2cta$m.err
  python -c "import json;d=json.load(open('$O/r2b_bench_2cta$m.json'));print('2CTA=$m ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'])" || tail -5 $O/r2b_bench_2cta$m.err
done
