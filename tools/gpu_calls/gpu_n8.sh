#!/bin/bash
# 8-GPU call: N=8 bench with both gather modes (weak scaling, 64 utterances per rank)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
N=${NGPU:-8}
for g in push nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 20 --warmup 3 --gather $g > $O/n${N}_bench_$g.json 2> $O/n${N}_bench_$g.err
  python -c "import json;d=json.loads(open('$O/n${N}_bench_$g.json').read().strip().splitlines()[-1]);print('$g N=$N value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])" || tail -15 $O/n${N}_bench_$g.err
done
