#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "busy" (exit 3, nothing charged)
# the built .so travels with the snapshot: make sure it is current first
python -c "import sys; sys.path.insert(0, '.'); from esmdiff_b200 import _lib; _lib.build()" || exit 1
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null
    sleep 90
done
exit 3
