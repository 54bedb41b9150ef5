#!/bin/bash
# GPU call 27 of round 2: first conv alone at the idle clock: stage-skipping variants of the row-streaming kernel, and the im2col kernel.
set -u
mkdir -p gpurun_out
L=reve_b200
O=gpurun_out/r02_c27_conv0_alone.txt
: > $O
for lib in libreve_cuda libreve_cuda_xp1 libreve_cuda_xp2 libreve_cuda_xp4 libreve_cuda_xp8 libreve_cuda; do
  REVE_LIB=$PWD/$L/$lib.so timeout 120 python tools/time_conv0.py >> $O 2>&1
done
timeout 120 python tools/time_conv0.py 1920x1080x2 128 >> $O 2>&1
cat $O
