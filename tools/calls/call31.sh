#!/bin/bash
# GPU call 31 of round 2: first conv alone at several frame sizes: fixed cost vs per-pixel cost, both kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c31_conv0_sizes.txt
: > $O
for sz in 64x64x2 480x270x2 960x540x2 1920x1080x2 3840x2160x2; do
  timeout 120 python tools/time_conv0.py $sz >> $O 2>&1
  timeout 120 python tools/time_conv0.py $sz 128 >> $O 2>&1
done
cut -c1-170 $O
