#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   - retries while the pod answers "busy" (nothing charged), up to ~40 min
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if ! grep -q "nothing was charged" "$log"; then exit $rc; fi
  sleep 90
done
exit 3
