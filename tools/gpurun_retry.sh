#!/bin/bash
# gpurun with retries while the pod answers "busy" (status=transient, nothing charged):
#   tools/gpurun_retry.sh LOGFILE [gpurun options] -- 'command'
# The snapshot of /root/repo is taken when a call is accepted, so keep the tree runnable while this waits.
LOG=$1; shift
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then exit 0; fi
  sleep 45
done
exit 3
