#!/bin/bash
set -u
mkdir -p gpurun_out
for c in 0 17; do
  REVE_DEBUG_TRACE=1 REVE_DEBUG_TRACE_CHAIN=$c REVE_CHAIN=4 TRACE_TOP=3 timeout 120 python tools/gpu_trace_chain.py 2>&1 | head -5 >> gpurun_out/r02_c20_chain_epilogue.txt
done
echo done
