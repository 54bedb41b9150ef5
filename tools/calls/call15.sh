#!/bin/bash
# GPU call 15 of round 2: publish cadence with the quarter hand-over (every row / every 2nd (shipped) / every 3rd).
set -u
mkdir -p gpurun_out
timeout 400 python tools/ab_libs.py reve_b200/libreve_cuda.so reve_b200/libreve_cuda_pub1.so reve_b200/libreve_cuda_pub3.so > gpurun_out/r02_c15_ab_publish_every.txt 2>&1
echo done
