#!/bin/bash
# STFT: vectorised up-front staging + MUFU log; tests + A/B timing of the three occupancy shapes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_frontend_gpu.py -m gpu -q -k "stft or frontend" > $O/r3b_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3b_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/r3b_tests.log | tail -12
for k in 0 1 2 0 1 2; do
  echo "STFT_OCC=$k" >> $O/r3b_stft.txt
  STYLER_STFT_OCC=$k timeout 120 python tools/prof_kernels.py --only stft_mel_c4 --iters 15 >> $O/r3b_stft.txt 2>&1
done
cat $O/r3b_stft.txt
