#!/usr/bin/env bash
# gpurun with retries on "no slot right now" (exit 3 / status=transient): tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'
# Development helper only (not part of the product).
T="$1"; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun_retry: gave up"; exit 3
