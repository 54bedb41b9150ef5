#!/bin/bash
# GPU call 33 of round 2: same-box, in-pipeline A/B of the two first-conv kernels (1080p x2, 720p x4, 540p x3), then the first-conv tests.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c33_ab_conv0.txt
: > $O
for sz in 1920x1080x2 1280x720x4 960x540x3; do
  for i in 1 2; do
    AB_SIZE=$sz REVE_DEBUG_FLAGS=128 timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda.so 2>&1 | head -1 | sed 's/^/{"conv0": "im2col", "r": /; s/$/}/' >> $O
    AB_SIZE=$sz timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda.so 2>&1 | head -1 | sed 's/^/{"conv0": "rows", "r": /; s/$/}/' >> $O
  done
done
cat $O
timeout 900 python -m pytest tests -x -q -m gpu -k "first_conv" 2>&1 | tail -2
