#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <timeout> <command...> ; retries while the pod answers "transient"
out=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$out" 2>&1
  if grep -q "status=transient\|busy" "$out" && ! grep -q "status=ok" "$out"; then sleep 150; continue; fi
  break
done
