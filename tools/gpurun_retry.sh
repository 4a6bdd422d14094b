#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...> — retries while the pod answers "transient" (nothing charged)
log="$1"; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient" "$log" || grep -q "no box or slot" "$log"; then
    sleep 45
    continue
  fi
  break
done
