#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/pytest_gpu.log | tail -12
timeout 200 python tools/prof_kernels.py --only bilstm_h80 2>&1 | tail -1
for P in 0 1; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --pipeline $P > $O/r2r_bench_p$P.json 2>$O/r2r_bench_p$P.err
  python -c "import json;d=json.load(open('$O/r2r_bench_p$P.json'));print('pipeline $P: value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'], d['clocks'])" || tail -5 $O/r2r_bench_p$P.err
done
STYLER_PIPE_PRIO=0 timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --pipeline 1 > $O/r2r_bench_p1prio0.json 2>$O/r2r_bench_p1prio0.err
python -c "import json;d=json.load(open('$O/r2r_bench_p1prio0.json'));print('pipeline 1 equal priority: value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'])"
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --pipeline 1 --precision fp16 > $O/r2r_bench_fp16.json 2>$O/r2r_bench_fp16.err
python -c "import json;d=json.load(open('$O/r2r_bench_fp16.json'));print('fp16 pipeline 1: value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'])"
timeout 300 python tools/timeline.py --csv $O/r2r_timeline.csv > $O/r2r_timeline.txt 2>&1; head -1 $O/r2r_timeline.txt
