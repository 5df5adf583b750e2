#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/pytest_gpu.log | tail -12
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras > $O/r2y_bench.json 2>$O/r2y_bench.err
python -c "import json;d=json.load(open('$O/r2y_bench.json'));print('value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'])" || tail -5 $O/r2y_bench.err
timeout 100 python tools/timeline.py --csv $O/r2y_timeline.csv > $O/r2y_timeline.txt 2>&1; head -1 $O/r2y_timeline.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py tests/test_loss_gpu.py -m gpu -q -x -k "tensor_core_recurrence or gn_calibrator or loss_vs_reference or bucket" > $O/r2y_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/r2y_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tensor_core_recurrence and bf16 and 18" > $O/r2y_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/r2y_racecheck.log
