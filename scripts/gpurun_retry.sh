#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers "busy" (exit 3, nothing charged)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "done rc=$rc try=$i" >> "$log"; exit $rc; fi
  sleep 120
done
echo "gave up" >> "$log"
