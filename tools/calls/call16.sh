#!/bin/bash
# GPU call 16 of round 2: senders ship their own quarters, the courier only publishes: parity, sanitizers, A/B, soak.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider -x -k "per_layer or golden or random_geometries or full_size or race_free or two_contexts or staged_path" 2>&1 | tail -15 ) > $O/r02_c16_pytest.log
timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda_quarters.so reve_b200/libreve_cuda.so > $O/r02_c16_ab_self_store.txt 2>&1
for tool in memcheck synccheck racecheck; do
  echo "== $tool (chains of 4 forced, grid capped at 8 CTAs)" >> $O/r02_c16_sanitize.txt
  REVE_CHAIN=4 REVE_DEBUG_GRID=8 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -vE "^=========\s*$" | tail -8 >> $O/r02_c16_sanitize.txt
done
timeout 600 python tools/race_hunt.py 1500 > $O/r02_c16_race_hunt_1500.txt 2>&1
for c in 0 17; do
  REVE_DEBUG_TRACE=1 REVE_DEBUG_TRACE_CHAIN=$c REVE_CHAIN=4 TRACE_TOP=3 timeout 120 python tools/gpu_trace_chain.py 2>&1 | head -5 >> $O/r02_c16_chain_waits.txt
done
echo done
