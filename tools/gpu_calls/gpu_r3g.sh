#!/bin/bash
# STFT direct-load variant (STFT_OCC=3): tests, racecheck, A/B timing against the staged persistent shape
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_frontend_gpu.py -m gpu -q -k "stft or frontend" > $O/r3g_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3g_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/r3g_tests.log | tail -12
for k in 1 3 1 3; do
  echo "STFT_OCC=$k" >> $O/r3g_stft.txt
  STYLER_STFT_OCC=$k timeout 100 python tools/prof_kernels.py --only stft_mel_c4 --iters 15 >> $O/r3g_stft.txt 2>&1
done
cat $O/r3g_stft.txt
STYLER_STFT_OCC=3 timeout 100 python -m pytest tests/test_frontend_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "stft_mel or frontend" > $O/r3g_tests_occ3.log 2>&1; echo "occ3 default pytest rc=$?"; tail -1 $O/r3g_tests_occ3.log
timeout 120 compute-sanitizer --tool racecheck python tools/stft_small.py > $O/r3g_stft_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 $O/r3g_stft_racecheck.log
