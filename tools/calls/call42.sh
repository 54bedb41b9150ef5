#!/bin/bash
# GPU call 42 of round 2: 640x480 x2 (canvas 723 columns): chains of 4 need 7 strips of 120, chains of 2 need 6 strips of 124.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c42_480p_chain_len.txt
: > $O
for i in 1 2; do
  for c in 4 2 1; do
    AB_SIZE=640x480x2 REVE_CHAIN=$c timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda.so 2>&1 | head -1 | sed "s/^/{\"layers_per_launch\": $c, \"r\": /; s/$/}/" >> $O
  done
done
cat $O
