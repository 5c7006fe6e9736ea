#!/bin/bash
# usage: tools/gpu/submit.sh <timeout_s> <script under tools/gpu> [gpus]   -- retries while the pod answers "transient"/busy
T=$1; S=$2; N=${3:-1}
name=$(basename "$S" .sh)
for attempt in $(seq 1 30); do
  if [ "$N" = "1" ]; then
    out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "bash $S > gpurun_out/$name.log 2>&1; tail -5 gpurun_out/$name.log" 2>&1)
  else
    out=$(/usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "bash $S > gpurun_out/$name.log 2>&1; tail -5 gpurun_out/$name.log" 2>&1)
  fi
  echo "$out" | tail -12
  if echo "$out" | grep -q "status=transient\|no box\|busy"; then
    echo "[submit] attempt $attempt: not scheduled, retrying in 150 s"; sleep 150; continue
  fi
  break
done
