#!/bin/bash
# gpurun_retry.sh <timeout-seconds> <logfile> <command...>: re-submit while the pod answers "busy / draining" (rc 3)
# GPUS=N in the environment asks for an N-GPU box (charged N x)
T=$1; LOG=$2; shift 2
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$T" -- "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|status=busy\|no box\|retry in a few minutes" "$LOG" && ! grep -q "status=ok" "$LOG"; then
    sleep 90; continue
  fi
  break
done
