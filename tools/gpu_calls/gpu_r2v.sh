#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/pytest_gpu.log | tail -12
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras > $O/r2v_bench.json 2>$O/r2v_bench.err
python -c "import json;d=json.load(open('$O/r2v_bench.json'));print('value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'])" || tail -5 $O/r2v_bench.err
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --pipeline 1 > $O/r2v_bench_p1.json 2>$O/r2v_bench_p1.err
python -c "import json;d=json.load(open('$O/r2v_bench_p1.json'));print('p1 value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'])" || tail -5 $O/r2v_bench_p1.err
timeout 300 python tools/timeline.py --csv $O/r2v_timeline.csv > $O/r2v_timeline.txt 2>&1; head -1 $O/r2v_timeline.txt
