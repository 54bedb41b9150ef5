#!/bin/bash
# GPU call 26 of round 2: which stage bounds the row-streaming first conv (variants that skip one stage each; wrong output, timing only).
set -u
mkdir -p gpurun_out
L=reve_b200
timeout 900 python tools/ab_libs.py $L/libreve_cuda.so $L/libreve_cuda_xp1.so $L/libreve_cuda_xp2.so $L/libreve_cuda_xp4.so $L/libreve_cuda_xp8.so > gpurun_out/r02_c26_conv0_stages.txt 2>&1
cat gpurun_out/r02_c26_conv0_stages.txt
