#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention or cta_pairs or bilstm or conv1d" > $O/r2c_new.log 2>&1; echo "new rc=$?" >> $O/r2c_new.log
grep -E "passed|failed|FAILED|Error|rc=|watchdog" $O/r2c_new.log | tail -15
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  STYLER_TC_2CTA=$1 STYLER_ATTN_PERSIST=$2 timeout 200 python tools/prof_kernels.py --B 128 --only ffn1,postnet1,ffn2_ln,fc_ln,pred_conv_ln,attention_qkv,attention_lowvar,bilstm_h80 > $O/r2c_prof_$1$2.txt 2>&1
  echo "== 2CTA=$1 ATTN_PERSIST=$2 (B=128)"; cat $O/r2c_prof_$1$2.txt
done
timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > $O/r2c_bench.json 2>$O/r2c_bench.err
python -c "import json;d=json.load(open('$O/r2c_bench.json'));print('bench ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'])" || tail -5 $O/r2c_bench.err
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/pytest_gpu.log | tail -30
