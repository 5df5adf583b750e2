#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_peer_gather_gpu.py tests/test_inspection_gpu.py -m gpu -q -x > $O/n2c_tests.log 2>&1; echo "rc=$?" >> $O/n2c_tests.log
grep -E "passed|failed|FAILED|Error|rc=|MISMATCH|PEER" $O/n2c_tests.log | tail -12
for P in 0 1; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 --pipeline $P > $O/n2c_bench_p$P.json 2> $O/n2c_bench_p$P.err
  python -c "import json;d=json.loads(open('$O/n2c_bench_p$P.json').read().strip().splitlines()[-1]);print('pipeline $P N=2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])" || tail -15 $O/n2c_bench_p$P.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 5 --warmup 1 --impl reference > $O/n2c_bench_ref.json 2> $O/n2c_bench_ref.err; tail -c 600 $O/n2c_bench_ref.json
