#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -m gpu -q -x -k "bilstm or attention or golden or bench_shape or pipelined" > $O/r2p_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2p_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/r2p_tests.log | tail -8
for POLY in 0 2 3 4; do
  echo "ATTN_POLY=$POLY"; STYLER_ATTN_POLY=$POLY timeout 200 python tools/prof_kernels.py --B 128 --only attention_qkv,attention_lowvar 2>&1 | tail -2
done
timeout 200 python tools/prof_kernels.py --only bilstm_h80 2>&1 | tail -1
STYLER_LSTM_MULTI=0 timeout 200 python tools/prof_kernels.py --only bilstm_h80 2>&1 | tail -1
for P in 0 1; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --pipeline $P > $O/r2p_bench_p$P.json 2>$O/r2p_bench_p$P.err
  python -c "import json;d=json.load(open('$O/r2p_bench_p$P.json'));print('pipeline $P: value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'], d['clocks'])" || tail -5 $O/r2p_bench_p$P.err
done
STYLER_ATTN_POLY=0 timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --pipeline 0 > $O/r2p_bench_poly0.json 2>$O/r2p_bench_poly0.err
python -c "import json;d=json.load(open('$O/r2p_bench_poly0.json'));print('poly0: value ms',d['ms_per_step'])"
timeout 300 python tools/timeline.py --csv $O/r2p_timeline.csv > $O/r2p_timeline.txt 2>&1; head -1 $O/r2p_timeline.txt
