#!/bin/bash
# GPU call 14 of round 2: the quarter hand-over as shipped: whole suite, sanitizers, determinism soak, bench, ncu summary.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 ) > $O/r02_c14_pytest.log
for tool in memcheck synccheck racecheck; do
  echo "== $tool (chains of 4 forced, grid capped at 8 CTAs)" >> $O/r02_c14_sanitize.txt
  REVE_CHAIN=4 REVE_DEBUG_GRID=8 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -vE "^=========\s*$" | tail -8 >> $O/r02_c14_sanitize.txt
done
timeout 600 python tools/race_hunt.py 3000 > $O/r02_c14_race_hunt_3000.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_c14_bench.json 2> $O/r02_c14_bench.err
timeout 300 python bench.py --workload 720p_x4 --no-cpu > $O/r02_c14_bench_720p.json 2> $O/r02_c14_bench_720p.err
timeout 300 python bench.py --workload 540p_x3 --no-cpu > $O/r02_c14_bench_540p.json 2> $O/r02_c14_bench_540p.err
timeout 300 python bench.py --workload 480p_x2 --no-cpu > $O/r02_c14_bench_480p.json 2> $O/r02_c14_bench_480p.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain -s 8 -c 1 -o $O/r02_c14_chain python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c14_ncu_chain.out 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_c14_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c14_launches.out 2>&1
echo done
