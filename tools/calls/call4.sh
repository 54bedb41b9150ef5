#!/bin/bash
# GPU call 4 of round 2 (2 GPUs): the single-process multi-GPU mode, torchrun with either process group, the C++
# scheduler over two devices, the config-5 stand-in on two devices, and the release-`consumed` A/B.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --single-process --gpus 2 --no-cpu > $O/r02_c4_bench_single_n2.json 2> $O/r02_c4_bench_single_n2.err
timeout 300 python bench.py --gpus 2 --no-cpu > $O/r02_c4_bench_torchrun_n2.json 2> $O/r02_c4_bench_torchrun_n2.err
timeout 300 python bench.py --gpus 2 --pg gloo --no-cpu > $O/r02_c4_bench_torchrun_gloo_n2.json 2> $O/r02_c4_bench_torchrun_gloo_n2.err
( timeout 300 python -m pytest tests/test_host_driver.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 ) > $O/r02_c4_pytest_driver.log
timeout 300 python tools/bench_e2e.py --gpus 2 --frames 240 > $O/r02_c4_e2e_n2.txt 2>&1
timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda.so reve_b200/libreve_cuda_consrel.so > $O/r02_c4_ab_consumed_release.txt 2>&1
echo done
