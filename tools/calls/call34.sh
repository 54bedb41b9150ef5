#!/bin/bash
# GPU call 34 of round 2 (run again as call 37 after the two-rows-per-step change): HEAD with the row-streaming first conv: whole suite, sanitizers, smoke, bench at every workload, launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 ) > $O/r02_c34_pytest.log; cat $O/r02_c34_pytest.log
: > $O/r02_c34_sanitize.txt
for tool in memcheck synccheck racecheck; do
  echo "== $tool (chains of 4 forced, grid capped at 8 CTAs)" >> $O/r02_c34_sanitize.txt
  REVE_CHAIN=4 REVE_DEBUG_GRID=8 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -vE "^=========\s*$" | tail -8 >> $O/r02_c34_sanitize.txt
done
tail -30 $O/r02_c34_sanitize.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_c34_smoke.txt 2>&1; tail -2 $O/r02_c34_smoke.txt
timeout 600 python bench.py > $O/r02_c34_bench.json 2> $O/r02_c34_bench.err
for wl in 720p_x4 540p_x3 480p_x2; do timeout 300 python bench.py --workload $wl --no-cpu > $O/r02_c34_bench_$wl.json 2> $O/r02_c34_bench_$wl.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_c34_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c34_launches.out 2>&1
cut -c1-300 $O/r02_c34_bench.json
