#!/bin/bash
# round-2 GPU call A: parity suite, bench lines (headline + configs[1] + configs[3]), ncu summaries for STFT / LR / BiLSTM
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
tail -5 $O/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"
tail -c 3000 $O/r2a_bench.json; tail -5 $O/r2a_bench.err
timeout 200 python bench.py --steps 20 --no-graph --no-extras --no-cpu-baseline > $O/r2a_bench_nograph.json 2>> $O/r2a_bench.err
timeout 200 python bench.py --workload fftblock --steps 50 > $O/r2a_bench_fftblock.json 2>> $O/r2a_bench.err
timeout 200 python bench.py --workload stft --steps 30 > $O/r2a_bench_stft.json 2>> $O/r2a_bench.err
cat $O/r2a_bench_fftblock.json $O/r2a_bench_stft.json | cut -c1-1200
timeout 200 python tools/prof_kernels.py > $O/r2a_prof_kernels.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_mel_kernel -s 2 -c 1 -o $O/r2a_stft -f python tools/prof_kernels.py --only stft_mel_c4 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bilstm_kernel -s 2 -c 1 -o $O/r2a_bilstm -f python tools/prof_kernels.py --only bilstm_h80 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lr_expand_kernel -s 2 -c 1 -o $O/r2a_lr -f python tools/prof_kernels.py --only length_regulator --iters 1 > /dev/null 2>&1
ls -la $O | tail -20
