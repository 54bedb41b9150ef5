#!/bin/bash
# GPU call 9 of round 2: the shipped build, as the driver will run it: suite, smoke, bench (+ launch list, + a soak).
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 ) > $O/r02_c9_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_c9_smoke.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_c9_bench.json 2> $O/r02_c9_bench.err
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > $O/r02_c9_bench_soak.json 2> $O/r02_c9_bench_soak.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_c9_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c9_launches.out 2>&1
echo done
