#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda_notrace.so reve_b200/libreve_cuda.so > gpurun_out/r02_c21_ab_trace_code.txt 2>&1
echo done
