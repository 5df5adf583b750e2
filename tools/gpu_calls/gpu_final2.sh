#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py > $O/final_bench.json 2>$O/final_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/final_bench.json'));print('value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'],d['roofline']['frac'],{k:(v.get('ms_per_step',v.get('graph_ms')),v.get('clocks')) for k,v in d['extras'].items()},d['cpu_baseline']['value'])" || tail -5 $O/final_bench.err
timeout 300 python bench.py --precision fp16 --no-extras --no-cpu-baseline --steps 30 > $O/final_bench_fp16.json 2>$O/final_bench_fp16.err
python -c "import json;d=json.load(open('$O/final_bench_fp16.json'));print('fp16 value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d['clocks'])"
