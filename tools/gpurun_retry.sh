#!/bin/bash
# gpurun_retry.sh <log> <gpurun args...>: retry while the pod answers "transient / busy" (exit 3), at most 12 times, 2 minutes apart
log=$1; shift
for i in $(seq 1 12); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
