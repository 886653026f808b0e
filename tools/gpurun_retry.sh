#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy / no box" (exit code 3, nothing charged).
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'
T="$1"; shift
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$T" "$@"
    rc=$?
    if [ "$rc" != "3" ]; then exit $rc; fi
    sleep 90
done
exit 3
