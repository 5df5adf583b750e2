#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/pytest_gpu.log | tail -12
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2x_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2x_smoke.log
timeout 600 python bench.py > $O/r2x_bench.json 2>$O/r2x_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/r2x_bench.json'));print('value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'],d['roofline']['frac'],{k:(v.get('ms_per_step',v.get('graph_ms')),v.get('clocks')) for k,v in d['extras'].items()},d['cpu_baseline']['value'])" || tail -5 $O/r2x_bench.err
timeout 200 python bench.py --workload stft --steps 30 > $O/r2x_bench_stft.json 2>>$O/r2x_bench.err
timeout 200 python bench.py --workload fftblock --steps 50 > $O/r2x_bench_fftblock.json 2>>$O/r2x_bench.err
python -c "
import json
for f in ('stft','fftblock'):
    d=json.load(open('$O/r2x_bench_%s.json'%f)); print(f, d['ms_per_step'], d['value'], d['roofline']['frac'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None)"
