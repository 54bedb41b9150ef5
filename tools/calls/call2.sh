#!/bin/bash
# GPU call 2 of round 2: coalesced tail stores + self-balancing chain split: tests, same-box A/B, timelines, ncu captures.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $O/r02_c2_pytest.log
# same-box A/B/A/B: previous build vs this build, and this build with the equal split
timeout 300 python tools/ab_libs.py reve_b200/libreve_cuda_prev.so reve_b200/libreve_cuda.so > $O/r02_c2_ab_prev_new.txt 2>&1
REVE_DEBUG_FLAGS=64 timeout 200 python tools/ab_libs.py reve_b200/libreve_cuda.so > $O/r02_c2_ab_equal_split.txt 2>&1
REVE_DEBUG_TRACE=1 REVE_CHAIN=4 timeout 120 python tools/gpu_trace_chain.py > $O/r02_c2_chain_timeline_balanced.txt 2>&1
REVE_DEBUG_FLAGS=64 REVE_DEBUG_TRACE=1 REVE_CHAIN=4 timeout 120 python tools/gpu_trace_chain.py > $O/r02_c2_chain_timeline_equal.txt 2>&1
timeout 600 python bench.py > $O/r02_c2_bench.json 2> $O/r02_c2_bench.err
# ncu: the tensor counters on a kernel that is 100 % tensor-busy by construction, then the three kernels of the step
timeout 300 ncu --set full --clock-control none -s 2 -c 1 -o $O/r02_c2_busy192 tools/microbench/bin/umma_ncu_counters 192 100000 > $O/r02_c2_busy192.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain -s 8 -c 1 -o $O/r02_c2_chain python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c2_ncu_chain.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv0 -s 2 -c 1 -o $O/r02_c2_conv0 python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c2_ncu_conv0.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 2 -c 1 -o $O/r02_c2_tail python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c2_ncu_tail.out 2>&1
echo done
