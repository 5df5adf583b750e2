#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
N=${NGPU:-8}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 20 --warmup 3 > $O/n${N}b_bench.json 2> $O/n${N}b_bench.err
python -c "import json;d=json.loads(open('$O/n${N}b_bench.json').read().strip().splitlines()[-1]);print('N=$N value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['clocks'])" || tail -15 $O/n${N}b_bench.err
