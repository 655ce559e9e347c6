#!/bin/bash
# usage: tools/gpu_call.sh <script under tools/calls/> [gpurun timeout seconds] [gpus]
# runs the script on a B200 box through gpurun, retrying while the pod answers "busy" (exit 3, nothing charged)
S=$1; T=${2:-1500}; G=${3:-1}
EXTRA=""; [ "$G" != "1" ] && EXTRA="--gpus $G"
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T $EXTRA -- "bash $S"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
