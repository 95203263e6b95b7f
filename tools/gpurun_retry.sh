#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script under tools/calls/> [gpus]
# Retries while the pod answers "busy" (exit code 3: nothing charged).
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" -gt 1 ]; then /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S"; else /usr/local/graft/bin/gpurun --timeout $T -- "bash $S"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
