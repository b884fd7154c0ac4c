#!/bin/bash
# usage: [GPUS=2] tools/gpu_retry.sh LOG TIMEOUT 'command'  -- retries gpurun while the pod answers "transient" (no slot free, nothing charged)
LOG=$1; TO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|exit code 3\|rc=3" $LOG && ! grep -q "status=ok" $LOG; then sleep 75; continue; fi
  break
done
echo "done after $i attempt(s)" >> $LOG
