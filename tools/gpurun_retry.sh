#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <timeout> [--gpus N] <command...> ; retries while the pod answers "transient"/busy
out=$1; shift; to=$1; shift
extra=()
if [ "$1" = "--gpus" ]; then extra=(--gpus "$2"); shift 2; fi
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > "$out" 2>&1
  if grep -q "status=transient\|busy" "$out" && ! grep -q "status=ok" "$out"; then sleep 150; continue; fi
  break
done
