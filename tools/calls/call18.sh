#!/bin/bash
# GPU call 18 of round 2: quarter stores in conv0 and in the single-layer body kernel: whole suite, A/B (default and
# single-layer mode), sanitizers.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 ) > $O/r02_c18_pytest.log
timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda_c0old.so reve_b200/libreve_cuda.so > $O/r02_c18_ab_conv0_quarters.txt 2>&1
REVE_SHARED_DEVICE=1 timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda_c0old.so reve_b200/libreve_cuda.so > $O/r02_c18_ab_single_layer_quarters.txt 2>&1
for tool in memcheck synccheck racecheck; do
  echo "== $tool (single-layer launches)" >> $O/r02_c18_sanitize.txt
  REVE_CHAIN=0 REVE_DEBUG_GRID=8 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -vE "^=========\s*$" | tail -6 >> $O/r02_c18_sanitize.txt
  echo "== $tool (CTA pairs)" >> $O/r02_c18_sanitize.txt
  REVE_CTA_PAIRS=1 REVE_DEBUG_GRID=8 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -vE "^=========\s*$" | tail -6 >> $O/r02_c18_sanitize.txt
done
echo done
