#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
N=8
nvidia-smi topo -m > $O/n8_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > $O/n8_lscpu.txt 2>&1
cat /sys/fs/cgroup/cpu.max >> $O/n8_lscpu.txt 2>&1; cat /sys/fs/cgroup/cpuset.cpus.effective >> $O/n8_lscpu.txt 2>&1
for d in noh2d nod2h; do
  STYLER_BENCH_E2E_DIAG=$d timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --steps 20 --warmup 3 --no-extras > $O/n8_diag_$d.json 2> $O/n8_diag_$d.err
  python -c "import json;d=json.loads(open('$O/n8_diag_$d.json').read().strip().splitlines()[-1]);print('$d N=8 value ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],d.get('cpu_binding'))" || tail -15 $O/n8_diag_$d.err
done
cat $O/n8_lscpu.txt; head -14 $O/n8_topo.txt
