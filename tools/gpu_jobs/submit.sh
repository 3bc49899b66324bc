#!/bin/bash
# usage: submit.sh <timeout-seconds> <job-script> [gpus]   -- retries while the pod answers "busy / draining"
t=$1; job=$2; gpus=${3:-1}
for i in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $t -- "bash $job" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $gpus --timeout $t -- "bash $job" 2>&1); fi
  echo "$out" | tail -15
  if echo "$out" | grep -q "status=transient\|rc=3\|retry in a few minutes"; then sleep 90; continue; fi
  break
done
