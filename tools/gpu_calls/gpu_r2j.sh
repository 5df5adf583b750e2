#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 200 python tools/prof_kernels.py --B 128 --only postnet_last,postnet0,qkv_fused,fc_ln,audio_c320_k5 2>&1 | tail -6
echo "== qkv as CTA pairs (TC_PERSIST=0 TC_2CTA=2)"
STYLER_TC_PERSIST=0 STYLER_TC_2CTA=2 timeout 200 python tools/prof_kernels.py --B 128 --only qkv_fused,fc_ln 2>&1 | tail -3
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -m gpu -q -x > $O/r2j_tests.log 2>&1; echo "rc=$?" >> $O/r2j_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/r2j_tests.log | tail -8
for i in 1 2; do
timeout 300 python bench.py --steps 20 --no-extras --no-cpu-baseline > $O/r2j_bench.json 2>$O/r2j_bench.err
python -c "import json;d=json.load(open('$O/r2j_bench.json'));print('bench ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'])" || tail -5 $O/r2j_bench.err
done
