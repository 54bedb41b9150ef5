#!/bin/bash
# GPU call 41 of round 2 (4 GPUs): HEAD (row-streaming first conv) under torchrun at N = 1, 2, 4 and as one process with four threads.
set -u
mkdir -p gpurun_out
O=gpurun_out
for n in 1 2 4; do
  timeout 400 python bench.py --gpus $n --steps 20 --warmup 5 --no-cpu > $O/r02_c41_bench_torchrun_n$n.json 2> $O/r02_c41_bench_torchrun_n$n.err
done
timeout 400 python bench.py --gpus 4 --single-process --no-cpu > $O/r02_c41_bench_single_n4.json 2> $O/r02_c41_bench_single_n4.err
echo done
