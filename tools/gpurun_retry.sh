#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <script>  -- retries while the pod answers "transient" (busy, nothing charged)
T=$1; shift
for i in $(seq 1 20); do
  OUT=$(gpurun --timeout $T -- "$@" 2>&1)
  echo "$OUT" | tail -150
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] pod busy, sleeping 150 s"; sleep 150; continue; fi
  break
done
