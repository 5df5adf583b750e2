#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/pytest_gpu.log | tail -30
timeout 200 python tools/prof_kernels.py --B 128 --only postnet_last,postnet0 2>&1 | tail -3
timeout 300 python bench.py --steps 20 > $O/r2i_bench.json 2>$O/r2i_bench.err
python -c "import json;d=json.load(open('$O/r2i_bench.json'));print('bench ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'],d['extras'])" || tail -5 $O/r2i_bench.err
