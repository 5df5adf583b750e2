#!/bin/bash
# final verification of this session's tree: full GPU suite, smoke, STFT racecheck, configs[3] bench line + ncu, headline bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/r3_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r3_pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/r3_pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r3_smoke.log
timeout 150 compute-sanitizer --tool racecheck python tools/stft_small.py > $O/r3_stft_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/r3_stft_racecheck.log
timeout 150 python bench.py --workload stft --steps 40 --warmup 5 > $O/r3_bench_stft.json 2> $O/r3_bench_stft.err
python -c "import json;d=json.loads(open('$O/r3_bench_stft.json').read().strip().splitlines()[-1]);print('bench stft',d['ms_per_step'],d['value'],d['roofline']['frac'],d['config']['launch'],d['e2e']['value'],d['clocks'])" || tail -5 $O/r3_bench_stft.err
timeout 150 ncu --set full --clock-control none --import-source on -k regex:stft_mel_kernel -s 2 -c 1 -o $O/r3_stft_final -f python tools/prof_kernels.py --only stft_mel_c4 --iters 1 > $O/r3_ncu.log 2>&1; ls -la $O/r3_stft_final.ncu-rep
timeout 300 python bench.py > $O/r3_bench.json 2>$O/r3_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads(open('$O/r3_bench.json').read().strip().splitlines()[-1]);print('value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'],d['roofline']['frac'])" || tail -5 $O/r3_bench.err
