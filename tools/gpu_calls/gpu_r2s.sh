#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 30 --warmup 3 > $O/r2s_bench.json 2> $O/r2s_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/r2s_bench.json'));print('value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'], d['roofline']['frac'], {k:v.get('ms_per_step', v) for k,v in d['extras'].items()}, d['cpu_baseline'])" || tail -5 $O/r2s_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sb:: -s 780 -c 270 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-extras --no-cpu-baseline > $O/r2s_launchbench.log 2>&1
wc -l $O/launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bilstm_mma_kernel -s 2 -c 1 -o $O/r2s_bilstm_mma -f python tools/prof_kernels.py --only bilstm_h80 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -o $O/r2s_attn -f python tools/prof_kernels.py --B 128 --only attention_lowvar --iters 1 > /dev/null 2>&1
timeout 200 python tools/prof_kernels.py --B 128 > $O/r2s_prof_kernels_b128.txt 2>&1; tail -25 $O/r2s_prof_kernels_b128.txt
ls -la $O/r2s*
