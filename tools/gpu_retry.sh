#!/bin/bash
# tools/gpu_retry.sh LOGFILE TIMEOUT CMD...  -- resubmit a gpurun call while the pod answers busy (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
