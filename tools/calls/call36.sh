#!/bin/bash
# GPU call 36 of round 2: row-streaming first conv with two rows per step: parity, alone, and in the pipeline against one row per step.
set -u
mkdir -p gpurun_out
L=$PWD/reve_b200
timeout 600 python tools/check_conv0_rows.py > gpurun_out/r02_c36_conv0_rows.txt 2> gpurun_out/r02_c36_conv0_rows.err; grep -c '"oracle_bad_frac": 0.0' gpurun_out/r02_c36_conv0_rows.txt; grep PARITY gpurun_out/r02_c36_conv0_rows.txt; tail -3 gpurun_out/r02_c36_conv0_rows.err
timeout 900 python -m pytest tests -x -q -m gpu -k "first_conv or per_layer" 2>&1 | tail -2
O=gpurun_out/r02_c36_conv0_alone.txt
: > $O
for i in 1 2; do for v in libreve_cuda_rows1 libreve_cuda; do REVE_LIB=$L/$v.so timeout 120 python tools/time_conv0.py >> $O 2>&1; done; done
cut -c1-170 $O
timeout 900 python tools/ab_libs.py $L/libreve_cuda_rows1.so $L/libreve_cuda.so > gpurun_out/r02_c36_ab.txt 2>&1; cat gpurun_out/r02_c36_ab.txt
