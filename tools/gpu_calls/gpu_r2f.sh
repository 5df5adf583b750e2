#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > $O/r2f_bench.json 2>$O/r2f_bench.err
python -c "import json;d=json.load(open('$O/r2f_bench.json'));print('bench ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'])" || tail -5 $O/r2f_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bilstm_kernel -s 2 -c 1 -o $O/r2f_bilstm -f python tools/prof_kernels.py --only bilstm_h80 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -o $O/r2f_attn -f python tools/prof_kernels.py --B 128 --only attention_lowvar --iters 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sb:: -s 700 -c 330 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-extras --no-cpu-baseline > $O/r2f_launchbench.log 2>&1
ls -la $O/r2f* $O/launches.csv | tail
