#!/bin/bash
# GPU call 30 of round 2: row-streaming first conv: producer / epilogue group counts (alone at the idle clock, then in the pipeline).
set -u
mkdir -p gpurun_out
L=$PWD/reve_b200
O=gpurun_out/r02_c30_conv0_groups.txt
: > $O
for i in 1 2; do
for v in p3e3 p3e4 p2e4 p2e5; do
  REVE_LIB=$L/libreve_cuda_$v.so timeout 120 python tools/time_conv0.py >> $O 2>&1
done
done
cat $O | cut -c1-160
timeout 900 python tools/ab_libs.py $L/libreve_cuda_p3e3.so $L/libreve_cuda_p3e4.so $L/libreve_cuda_p2e4.so $L/libreve_cuda_p2e5.so > gpurun_out/r02_c30_ab_groups.txt 2>&1
cat gpurun_out/r02_c30_ab_groups.txt
