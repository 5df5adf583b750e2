#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_peer_gather_gpu.py -m gpu -q -x > $O/n2_tests.log 2>&1; echo "rc=$?" >> $O/n2_tests.log
grep -E "passed|failed|FAILED|Error|rc=|MISMATCH|PEER" $O/n2_tests.log | tail -12
for g in push nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 --gather $g > $O/n2_bench_$g.json 2> $O/n2_bench_$g.err
  python -c "import json;d=json.loads(open('$O/n2_bench_$g.json').read().strip().splitlines()[-1]);print('$g N=2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d.get('cpu_binding'))" || tail -15 $O/n2_bench_$g.err
done
