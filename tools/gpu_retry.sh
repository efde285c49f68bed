#!/bin/sh
# tools/gpu_retry.sh <log> [gpurun args...] -- '<command>' : retries while the pod answers "transient" (exit 3)
log="$1"; shift
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 120
done
exit 3
