#!/bin/sh
# usage: [GPUS=N] tools/gpu/submit.sh <log> <timeout> <command...>   -- retries while the pod is busy (rc 3)
log=$1; shift; to=$1; shift
i=0
while [ $i -lt 40 ]; do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> $log; echo done >> $log; exit $rc; fi
  i=$((i+1)); sleep 90
done
echo "gave up" >> $log; echo done >> $log
