#!/bin/bash
# usage: scratch/gpu_calls/run.sh <script> <timeout> [gpus]   -- retries while the pod is busy (nothing is charged for those)
s=$1; t=$2; g=${3:-1}
log=gpurun_out/$(basename $s .sh).log
for i in $(seq 1 12); do
  if [ "$g" = "1" ]; then gpurun --timeout $t -- "bash $s" > $log 2>&1; else gpurun --gpus $g --timeout $t -- "bash $s" > $log 2>&1; fi
  if grep -q "status=transient\|status=busy\|rc=3" $log; then sleep 45; continue; fi
  break
done
