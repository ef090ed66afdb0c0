#!/bin/bash
# gpurun_retry_n.sh <gpus> <log> <timeout> <command...>: multi-GPU variant of gpurun_retry.sh
n=$1; shift; log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> $log; exit $rc; fi
  sleep 60
done
