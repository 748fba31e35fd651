#!/bin/bash
# usage: scripts/gpu_retry.sh <tag> <timeout> <command...> : retries while the pod answers busy (exit 3)
tag=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > /root/repo/gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /root/repo/gpurun_out/${tag}_call.log; then break; fi
  sleep 90
done
tail -3 /root/repo/gpurun_out/${tag}_call.log
