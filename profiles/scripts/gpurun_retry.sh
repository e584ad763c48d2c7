#!/bin/bash
# usage: gpurun_retry.sh <log> <gpus> <timeout> <command...>   -- retries while the pod answers "busy" (exit 3)
LOG=$1; GPUS=$2; TMO=$3; shift 3
for i in $(seq 1 40); do
  if [ "$GPUS" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1; else /usr/local/graft/bin/gpurun --gpus $GPUS --timeout $TMO -- "$@" > $LOG 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
