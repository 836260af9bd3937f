#!/bin/bash
# usage: tools/gpu.sh <logfile> <timeout_s> [--gpus N]   -- runs `bash tools/_call.sh` on a B200 box through gpurun, retrying while the pod is busy
log=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" -- 'bash tools/_call.sh' > $log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $log && ! grep -q "charged=[1-9]" $log; then sleep 90; continue; fi
  break
done
tail -60 $log
