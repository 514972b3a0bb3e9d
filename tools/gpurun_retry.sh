#!/bin/bash
# tools/gpurun_retry.sh LOG [gpurun args...] -- retries a gpurun call while the pod answers "busy" (exit code 3), at most 12 times
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 150
done
exit 3
