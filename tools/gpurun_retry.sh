#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout_s> <command...>   — retries while the pod answers "transient" / busy (nothing charged)
GPUS=$1; shift; TMO=$1; shift
for i in $(seq 1 20); do
  if [ "$GPUS" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $TMO -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $GPUS --timeout $TMO -- "$@" 2>&1); fi
  RC=$?
  if echo "$OUT" | grep -q "status=transient\|status=busy\|retry in a few minutes"; then echo "[retry $i] pod busy; sleeping"; sleep 150; continue; fi
  if [ $RC -eq 3 ]; then echo "[retry $i] rc=3; sleeping"; sleep 150; continue; fi
  echo "$OUT" | tail -100
  exit $RC
done
echo "gave up"
exit 3
