#!/bin/bash
# GPU call 32 of round 2: write-only bandwidth by store path with the first conv's access pattern.
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_paths tools/microbench/store_paths.cu -lcuda && timeout 120 /tmp/store_paths > gpurun_out/r02_c32_store_paths.txt 2>&1
cat gpurun_out/r02_c32_store_paths.txt
