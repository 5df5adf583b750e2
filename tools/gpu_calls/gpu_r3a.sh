#!/bin/bash
# STFT occupancy shapes: bitwise test + A/B timing (STFT_OCC = 0 | 1 | 2) at BASELINE configs[3]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_frontend_gpu.py -m gpu -q -k "stft or frontend" > $O/r3a_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3a_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/r3a_tests.log | tail -12
for k in 0 1 2 0 1 2; do
  echo "STFT_OCC=$k" >> $O/r3a_stft.txt
  STYLER_STFT_OCC=$k timeout 120 python tools/prof_kernels.py --only stft_mel_c4 --iters 15 >> $O/r3a_stft.txt 2>&1
done
cat $O/r3a_stft.txt
for k in 0 1 2; do
  STYLER_STFT_OCC=$k timeout 200 python bench.py --workload stft --steps 40 --warmup 5 --no-cpu-baseline > $O/r3a_bench_stft_occ$k.json 2> $O/r3a_bench_stft_occ$k.err
  python -c "import json;d=json.loads(open('$O/r3a_bench_stft_occ$k.json').read().strip().splitlines()[-1]);print('occ$k',d['ms_per_step'],d['value'],d['roofline']['frac'],d['e2e']['value'],d['clocks'])" || tail -5 $O/r3a_bench_stft_occ$k.err
done
