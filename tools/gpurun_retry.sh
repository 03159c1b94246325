#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <log> <command...>   -- retry while the pod answers busy/draining (rc 3, nothing charged)
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 150
done
exit 3
