#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <out-file> <command...>   -- retries while the pod answers "busy" (exit 3)
T=$1; OUT=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $OUT 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
