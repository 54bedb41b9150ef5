#!/bin/bash
# GPU call 25 of round 2: row-streaming first conv: 1, 2 or 3 output staging buffers per epilogue group, in the pipeline.
set -u
mkdir -p gpurun_out
L=reve_b200
timeout 900 python tools/ab_libs.py $L/libreve_cuda_ob1.so $L/libreve_cuda.so $L/libreve_cuda_ob3.so > gpurun_out/r02_c25_ab_outbufs.txt 2>&1
cat gpurun_out/r02_c25_ab_outbufs.txt
AB_SIZE=1280x720x4 timeout 900 python tools/ab_libs.py $L/libreve_cuda_ob1.so $L/libreve_cuda.so $L/libreve_cuda_ob3.so > gpurun_out/r02_c25_ab_outbufs_720.txt 2>&1
cat gpurun_out/r02_c25_ab_outbufs_720.txt
