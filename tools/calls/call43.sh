#!/bin/bash
# GPU call 43 of round 2: what the driver runs at round end, on HEAD: GPU suite, smoke, both bench arms.
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 ) > $O/r02_c43_pytest.log; cat $O/r02_c43_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_c43_smoke.txt 2>&1; tail -2 $O/r02_c43_smoke.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02_c43_bench_reference.json 2> $O/r02_c43_bench_reference.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_c43_bench.json 2> $O/r02_c43_bench.err
cut -c1-200 $O/r02_c43_bench_reference.json; cut -c1-200 $O/r02_c43_bench.json
