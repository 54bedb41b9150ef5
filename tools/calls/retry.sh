#!/bin/bash
# usage: tools/calls/retry.sh <gpus> <timeout> <script>   -- retries while the pod answers "busy / transient"
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --gpus $1 --timeout $2 -- "bash $3" 2>&1)
  echo "$out" | tail -6
  if echo "$out" | grep -q "status=ok"; then exit 0; fi
  if ! echo "$out" | grep -qE "status=transient|busy|retry"; then exit 1; fi
  sleep 240
done
exit 3
