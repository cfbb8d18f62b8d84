#!/bin/bash
# gpu_retry.sh LOG [gpurun options] -- 'command': runs gpurun, retrying while the pod answers "busy" (exit code 3)
log=$1; shift
for i in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 60
done
exit 3
