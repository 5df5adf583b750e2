#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_loss_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "loss or bilstm" > $O/r2u_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2u_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/r2u_tests.log | tail -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2u_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/r2u_smoke.log
timeout 200 python tools/prof_kernels.py --only bilstm_h80 2>&1 | tail -1
timeout 300 python bench.py --steps 30 --no-cpu-baseline > $O/r2u_bench.json 2>$O/r2u_bench.err
python -c "import json;d=json.load(open('$O/r2u_bench.json'));print('value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'], {k:v.get('ms_per_step', v) for k,v in d['extras'].items()})" || tail -5 $O/r2u_bench.err
