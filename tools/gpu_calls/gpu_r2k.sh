#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/pytest_gpu.log | tail -30
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/r2k_smoke.log
timeout 200 python tools/forward_phases.py 2>&1 | tail -5
timeout 200 python tools/prof_kernels.py --only bilstm_h80,groupnorm_relu 2>&1 | tail -3
# compute-sanitizer over the CUDA-core / glue kernels and one small tensor-core forward
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "length_regulator or bucket or bilstm or calibrator or quantize or classifier or onehot or stft or duration" > $O/r2k_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/r2k_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "bilstm or classifier or length_regulator or calibrator" > $O/r2k_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/r2k_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "tf_const_b2_l16 and bf16" > $O/r2k_memcheck_fwd.log 2>&1; echo "memcheck fwd rc=$?"; tail -3 $O/r2k_memcheck_fwd.log
