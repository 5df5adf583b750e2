#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout> <script>   -- retries while the pod answers "transient" (nothing charged)
G=""
if [ "$1" == "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; S=$2
for i in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun $G --timeout $T -- "bash $S" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|backing off"; then sleep 90; continue; fi
  echo "$OUT" | grep -v "^+" | tail -60
  exit 0
done
echo "gave up after 20 transient answers"
