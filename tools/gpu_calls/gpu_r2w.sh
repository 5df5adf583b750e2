#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
STYLER_BENCH_DEBUG=1 timeout -s INT 100 python -X faulthandler -u bench.py --steps 10 --no-cpu-baseline --no-extras --pipeline 1 > $O/r2w_bench_p1.json 2>$O/r2w_bench_p1.err; echo "p1 rc=$?"
tail -30 $O/r2w_bench_p1.err | cut -c1-300
python -c "import json;d=json.load(open('$O/r2w_bench_p1.json'));print('p1 value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'])"
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_inspection_gpu.py -m gpu -q -x > $O/r2w_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2w_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/r2w_tests.log | tail -8
timeout 200 python bench.py --steps 30 --no-cpu-baseline --no-extras > $O/r2w_bench.json 2>$O/r2w_bench.err
python -c "import json;d=json.load(open('$O/r2w_bench.json'));print('p0 value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'])" || tail -5 $O/r2w_bench.err
timeout 100 python tools/timeline.py --csv $O/r2w_timeline.csv > $O/r2w_timeline.txt 2>&1; head -1 $O/r2w_timeline.txt
