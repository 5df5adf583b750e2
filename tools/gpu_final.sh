#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " $O/pytest_gpu.log | tail -8
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.log
timeout 600 python bench.py > $O/final_bench.json 2>$O/final_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/final_bench.json'));print('value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'],d['roofline']['frac'],{k:(v.get('ms_per_step',v.get('graph_ms')),v.get('clocks')) for k,v in d['extras'].items()},d['cpu_baseline']['value'])" || tail -5 $O/final_bench.err
timeout 100 python tools/timeline.py > $O/final_timeline.txt 2>&1; head -1 $O/final_timeline.txt; grep -E "gn_calibrator|groupnorm" $O/final_timeline.txt | tail -3
