#!/bin/bash
# GPU call 1 of round 2: the full GPU test suite with the cooperative chained launch, the new bench protocol,
# the other workloads, a chain timeline and the ncu tensor-counter reconciliation.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/r02_c1_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 ) > $O/r02_c1_pytest.log
echo "pytest rc=$?" >> $O/r02_c1_pytest.log
timeout 600 python bench.py > $O/r02_c1_bench.json 2> $O/r02_c1_bench.err
timeout 300 python bench.py --workload 720p_x4 --no-cpu > $O/r02_c1_bench_720p.json 2> $O/r02_c1_bench_720p.err
timeout 300 python bench.py --workload 540p_x3 --no-cpu > $O/r02_c1_bench_540p.json 2> $O/r02_c1_bench_540p.err
timeout 300 python bench.py --shared-device --no-cpu --steps 10 > $O/r02_c1_bench_shared.json 2> $O/r02_c1_bench_shared.err
REVE_DEBUG_TRACE=1 REVE_CHAIN=4 timeout 120 python tools/gpu_trace_chain.py > $O/r02_c1_chain_timeline.txt 2>&1
M=sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for n in 192 256 64; do
  timeout 120 ncu --metrics $M --clock-control none -s 2 -c 1 --csv --log-file $O/r02_c1_counters_n$n.csv tools/microbench/bin/umma_ncu_counters $n 100000 > $O/r02_c1_counters_n$n.out 2>&1
done
timeout 60 ncu --query-metrics 2>/dev/null | grep -i tensor > $O/r02_c1_tensor_metrics.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_c1_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c1_launches.out 2>&1
echo done
