#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python tools/timeline.py --precision tf32 > $O/r2t_timeline_tf32.txt 2>&1; head -1 $O/r2t_timeline_tf32.txt; tail -26 $O/r2t_timeline_tf32.txt
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-extras --precision tf32 > $O/r2t_bench_tf32.json 2>$O/r2t_bench_tf32.err
python -c "import json;d=json.load(open('$O/r2t_bench_tf32.json'));print('tf32: value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'])" || tail -5 $O/r2t_bench_tf32.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 300 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-extras --no-cpu-baseline > $O/r2t_launchbench.log 2>&1
wc -l $O/launches.csv
