#!/bin/bash
# dev helper (build container): retry a gpurun call while the pod answers "busy" (exit code 3)
# usage: tools/gpu_retry.sh [gpurun options] -- 'command'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
