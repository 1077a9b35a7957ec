#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'  — re-submits while the pod answers busy/transient (nothing charged)
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|exit code 3\|no box"; then
    sleep 60
    continue
  fi
  echo "$out"
  exit 0
done
echo "gave up after 40 tries"; echo "$out" | tail -5
