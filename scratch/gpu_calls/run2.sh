#!/bin/bash
# usage: scratch/gpu_calls/run2.sh <script> <timeout> [gpus] [tries]  -- like run.sh, but keeps retrying for longer (busy pod, multi-GPU slots)
s=$1; t=$2; g=${3:-1}; n=${4:-80}
log=gpurun_out/$(basename $s .sh).log
for i in $(seq 1 $n); do
  if [ "$g" = "1" ]; then gpurun --timeout $t -- "bash $s" > $log 2>&1; else gpurun --gpus $g --timeout $t -- "bash $s" > $log 2>&1; fi
  if grep -q "status=transient\|status=busy\|rc=3\|backing off" $log; then sleep 60; continue; fi
  break
done
