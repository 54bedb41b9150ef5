#!/bin/bash
# GPU call 3 of round 2: tests after the fixes, tail output path A/B per scale, per-launch-kind balancing, stand-alone
# pack / unpack kernels, config-5 stand-in and single-process mode on one GPU.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $O/r02_c3_pytest.log
for sz in 1920x1080x2 960x540x3 1280x720x4; do
  AB_SIZE=$sz timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda_tail0.so reve_b200/libreve_cuda_tail28.so >> $O/r02_c3_ab_tail.txt 2>&1
done
timeout 200 python tools/ab_libs.py reve_b200/libreve_cuda_prev.so reve_b200/libreve_cuda.so > $O/r02_c3_ab_prev_new.txt 2>&1
REVE_DEBUG_TRACE=1 REVE_CHAIN=4 timeout 120 python tools/gpu_trace_chain.py > $O/r02_c3_chain_timeline_balanced.txt 2>&1
timeout 120 python tools/bench_pack.py > $O/r02_c3_bench_pack.txt 2>&1
timeout 300 python tools/bench_e2e.py --gpus 1 --frames 240 > $O/r02_c3_e2e_n1.txt 2>&1
timeout 300 python bench.py --single-process --gpus 1 --no-cpu --steps 10 > $O/r02_c3_bench_single_n1.json 2> $O/r02_c3_bench_single_n1.err
timeout 300 python bench.py --workload 480p_x2 --no-cpu > $O/r02_c3_bench_480p.json 2> $O/r02_c3_bench_480p.err
echo done
