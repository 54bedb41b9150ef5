#!/bin/bash
# GPU call 11 of round 2: what the MMA issuer of each layer of a chain waits for (input row vs accumulator slot).
set -u
mkdir -p gpurun_out
for c in 0 5 17 30; do
  REVE_DEBUG_TRACE=1 REVE_DEBUG_TRACE_CHAIN=$c REVE_CHAIN=4 TRACE_TOP=3 timeout 120 python tools/gpu_trace_chain.py 2>&1 | head -5 >> gpurun_out/r02_c11_chain_waits.txt
  echo "-- chain $c" >> gpurun_out/r02_c11_chain_waits.txt
done
echo done
