#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_frontend_gpu.py -m gpu -q -x -k "stft or front_end or classifier" > $O/r2l_tests.log 2>&1; echo "rc=$?" >> $O/r2l_tests.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/r2l_tests.log | tail -5
timeout 200 python tools/prof_kernels.py --only stft_mel_c4 2>&1 | tail -2
timeout 200 python bench.py --workload stft --steps 30 --no-cpu-baseline > $O/r2l_stft.json 2>$O/r2l_stft.err; python -c "import json;d=json.load(open('$O/r2l_stft.json'));print('stft ms',d['ms_per_step'],'frac',d['roofline']['frac'])" || tail -3 $O/r2l_stft.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_mel_kernel -s 2 -c 1 -o $O/r2l_stft -f python tools/prof_kernels.py --only stft_mel_c4 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:classifier_tail -s 1 -c 1 -o $O/r2l_cls -f python -m pytest tests/test_kernels_gpu.py -m gpu -q -k classifier > /dev/null 2>&1
ls -la $O/r2l*
