#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_forward_gpu.py tests/test_inspection_gpu.py -m gpu -q -s -k "golden or inspection or infer" > $O/r2z_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2z_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E |bucketize" $O/r2z_tests.log | tail -12
