#!/bin/bash
# GPU call 19 of round 2: HEAD as the driver will run it.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8 ) > $O/r02_c19_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_c19_smoke.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_c19_bench.json 2> $O/r02_c19_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_c19_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_c19_launches.out 2>&1
echo done
