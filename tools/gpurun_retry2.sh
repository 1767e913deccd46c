#!/bin/bash
# usage: tools/gpurun_retry2.sh <gpus> <timeout> <out-file> <command...>
G=$1; T=$2; OUT=$3; shift 3
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > $OUT 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
