#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit code 3): usage  scripts/gpurun_retry.sh <timeout> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 20); do
  if [ "$G" = 1 ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] pod busy, attempt $i"; sleep 90
done
exit 3
