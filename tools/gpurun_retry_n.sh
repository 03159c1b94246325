#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout-seconds> <log> <command...>
g=$1; t=$2; log=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 150
done
exit 3
