#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -m gpu -q -x -k "bilstm or golden or bench_shape or pipelined or audio" > $O/r2q_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2q_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/r2q_tests.log | tail -12
timeout 200 python tools/prof_kernels.py --only bilstm_h80 2>&1 | tail -1
for P in 0 1; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --pipeline $P > $O/r2q_bench_p$P.json 2>$O/r2q_bench_p$P.err
  python -c "import json;d=json.load(open('$O/r2q_bench_p$P.json'));print('pipeline $P: value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'], d['clocks'])" || tail -5 $O/r2q_bench_p$P.err
done
timeout 300 python tools/timeline.py --csv $O/r2q_timeline.csv > $O/r2q_timeline.txt 2>&1; head -1 $O/r2q_timeline.txt; tail -28 $O/r2q_timeline.txt
