#!/bin/bash
# GPU call 12 of round 2: hand-over staging in quarters (one TMA store per epilogue warp): parity, A/B, trace.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "per_layer or golden or random_geometries or full_size or race_free or two_contexts or staged_path" 2>&1 | tail -15 ) > $O/r02_c12_pytest.log
timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda_stage1.so reve_b200/libreve_cuda.so > $O/r02_c12_ab_quarters.txt 2>&1
for c in 0 17; do
  REVE_DEBUG_TRACE=1 REVE_DEBUG_TRACE_CHAIN=$c REVE_CHAIN=4 TRACE_TOP=3 timeout 120 python tools/gpu_trace_chain.py 2>&1 | head -5 >> $O/r02_c12_chain_waits_quarters.txt
done
echo done
