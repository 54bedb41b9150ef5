#!/bin/bash
# GPU call 44 of round 2 (8 GPUs): HEAD as one process with eight threads, and under torchrun.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --gpus 8 --single-process --no-cpu > $O/r02_c44_bench_single_n8.json 2> $O/r02_c44_bench_single_n8.err
timeout 300 python bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > $O/r02_c44_bench_torchrun_n8.json 2> $O/r02_c44_bench_torchrun_n8.err
cut -c1-160 $O/r02_c44_bench_single_n8.json; cut -c1-160 $O/r02_c44_bench_torchrun_n8.json
