#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> <gpus> <cmd...> ; retries while the pod answers "busy" (exit 3)
T=$1; shift; G=$1; shift
for i in $(seq 1 12); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
