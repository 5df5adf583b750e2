#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "bilstm or conv1d" > $O/r2e_new.log 2>&1; echo "new rc=$?" >> $O/r2e_new.log
grep -E "passed|failed|FAILED|Error|rc=|watchdog" $O/r2e_new.log | tail -15
for w in 0 1; do
STYLER_TC_WIDE=$w timeout 200 python tools/prof_kernels.py --only audio_c320_k5,bilstm_h80,fc_ln,pred_conv_ln > $O/r2e_prof_w$w.txt 2>&1
echo "== WIDE=$w"; cat $O/r2e_prof_w$w.txt
done
timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -x > $O/r2e_fwd.log 2>&1; echo "fwd rc=$?" >> $O/r2e_fwd.log
grep -E "passed|failed|FAILED|Error|rc=" $O/r2e_fwd.log | tail -8
timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > $O/r2e_bench.json 2>$O/r2e_bench.err
python -c "import json;d=json.load(open('$O/r2e_bench.json'));print('bench ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'])" || tail -5 $O/r2e_bench.err
