#!/bin/bash
# GPU call 24 of round 2: row-streaming first conv as the default: new tests, HBM ceilings by direction, A/B, full suite, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hbm_rw tools/microbench/hbm_rw.cu && /tmp/hbm_rw > $O/r02_c24_hbm_rw.txt 2>&1
cat $O/r02_c24_hbm_rw.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "first_conv" > $O/r02_c24_pytest_first_conv.txt 2>&1; tail -3 $O/r02_c24_pytest_first_conv.txt
timeout 600 python tools/check_conv0_rows.py > $O/r02_c24_conv0_rows.txt 2> $O/r02_c24_conv0_rows.err; tail -6 $O/r02_c24_conv0_rows.txt | cut -c1-150
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_c24_pytest_gpu.txt 2>&1; tail -3 $O/r02_c24_pytest_gpu.txt
timeout 600 python bench.py > $O/r02_c24_bench.json 2> $O/r02_c24_bench.err; cut -c1-400 $O/r02_c24_bench.json
