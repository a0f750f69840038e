#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free; nothing charged).  Usage: gpurun_retry.sh <timeout_s> <command...>
T=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_retry.$$ 2>&1
  if grep -q "status=transient" /tmp/gpurun_retry.$$; then sleep 150; continue; fi
  cat /tmp/gpurun_retry.$$; rm -f /tmp/gpurun_retry.$$; exit 0
done
cat /tmp/gpurun_retry.$$; rm -f /tmp/gpurun_retry.$$; echo "gave up after 12 transient answers"
