#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -q -x > $O/r2d_fwd.log 2>&1; echo "fwd rc=$?" >> $O/r2d_fwd.log
grep -E "passed|failed|FAILED|Error|rc=" $O/r2d_fwd.log | tail -8
timeout 200 python bench.py --steps 20 --no-extras --no-cpu-baseline > $O/r2d_bench.json 2>$O/r2d_bench.err
python -c "import json;d=json.load(open('$O/r2d_bench.json'));print('bench ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'roof',d['roofline']['avg_launch_ms'],d['roofline']['frac'])" || tail -5 $O/r2d_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bilstm_kernel -s 2 -c 1 -o $O/r2d_bilstm -f python tools/prof_kernels.py --only bilstm_h80 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -o $O/r2d_attn -f python tools/prof_kernels.py --B 128 --only attention_qkv --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -s 2 -c 1 -o $O/r2d_ffn2 -f python tools/prof_kernels.py --B 128 --only ffn2_ln --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -s 2 -c 1 -o $O/r2d_ffn1 -f python tools/prof_kernels.py --B 128 --only ffn1 --iters 1 > /dev/null 2>&1
ls -la $O/*.ncu-rep | tail -5
