#!/bin/bash
# gpurun with retries while the pod answers "no slot right now" (exit code 3: nothing charged)
#   tools/gpurun_retry.sh [--gpus N] TIMEOUT 'command'
G=()
if [ "$1" = "--gpus" ]; then G=(--gpus "$2"); shift 2; fi
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "${G[@]}" --timeout "$T" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  echo "[retry $i] no slot, sleeping 90 s"; sleep 90
done
exit 3
