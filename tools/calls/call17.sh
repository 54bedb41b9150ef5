#!/bin/bash
# GPU call 17 of round 2 (8 GPUs): the shipped build in the product's shape, configs[1] and configs[2], and under torchrun.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python bench.py --single-process --gpus 8 --no-cpu > $O/r02_c17_bench_single_n8.json 2> $O/r02_c17_bench_single_n8.err
timeout 400 python bench.py --single-process --gpus 8 --workload 720p_x4 --no-cpu > $O/r02_c17_bench_single_n8_720p.json 2> $O/r02_c17_bench_single_n8_720p.err
timeout 400 python bench.py --gpus 8 --no-cpu > $O/r02_c17_bench_torchrun_n8.json 2> $O/r02_c17_bench_torchrun_n8.err
echo done
