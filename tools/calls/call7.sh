#!/bin/bash
# GPU call 7 of round 2: waiting warps that sleep instead of spinning (conv0, tail) A/B; suite with the tightened
# tolerance; compute-sanitizer on the round-2 kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python tools/ab_libs.py reve_b200/libreve_cuda.so reve_b200/libreve_cuda_conv0_wait_sleep_32.so reve_b200/libreve_cuda_conv0_wait_sleep_200.so reve_b200/libreve_cuda_tail_wait_sleep_100.so > $O/r02_c7_ab_wait_sleep.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 ) > $O/r02_c7_pytest.log
for tool in memcheck synccheck racecheck; do
  echo "== $tool (chains of 4 forced, grid capped at 8 CTAs)" >> $O/r02_c7_sanitize.txt
  REVE_CHAIN=4 REVE_DEBUG_GRID=8 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -vE "^=========\s*$" | tail -12 >> $O/r02_c7_sanitize.txt
done
echo done
