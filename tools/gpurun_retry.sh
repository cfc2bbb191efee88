#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "busy" (exit 3, nothing charged)
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null
    sleep 90
done
exit 3
