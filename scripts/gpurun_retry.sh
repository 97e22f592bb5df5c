#!/bin/bash
# gpurun with retries while the pod answers busy (exit 3): scripts/gpurun_retry.sh LOG TIMEOUT -- 'command'
LOG=$1; TMO=$2; shift 3
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "nothing was charged" $LOG; then exit $rc; fi
  sleep 90
done
exit 3
