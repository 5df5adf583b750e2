#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_kernels_gpu.py tests/test_inspection_gpu.py -m gpu -q -x > $O/r2n_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2n_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=|Error|assert" $O/r2n_tests.log | tail -15
for P in 0 1; do
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --pipeline $P > $O/r2n_bench_p$P.json 2>$O/r2n_bench_p$P.err
  python -c "import json;d=json.load(open('$O/r2n_bench_p$P.json'));print('pipeline $P: value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'], d['clocks'])" || tail -5 $O/r2n_bench_p$P.err
done
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras --precision fp16 > $O/r2n_bench_fp16.json 2>$O/r2n_bench_fp16.err
python -c "import json;d=json.load(open('$O/r2n_bench_fp16.json'));print('fp16: value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'])" || tail -5 $O/r2n_bench_fp16.err
timeout 300 python tools/timeline.py --csv $O/r2n_timeline.csv > $O/r2n_timeline.txt 2>&1; head -1 $O/r2n_timeline.txt
