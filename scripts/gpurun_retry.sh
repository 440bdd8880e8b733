#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit code 3, nothing charged).  Usage: gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && break
  sleep 90
done
echo "gpurun_retry: rc=$rc after $i attempt(s)" >> "$log"
