#!/bin/bash
# GPU call 6 of round 2: feature error statistics (to set the per-layer tolerance from data), the suite, smoke, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/feature_errors.py > $O/r02_c6_feature_errors.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 ) > $O/r02_c6_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_c6_smoke.txt 2>&1
timeout 600 python bench.py > $O/r02_c6_bench.json 2> $O/r02_c6_bench.err
timeout 300 python bench.py --numa --no-cpu --steps 10 > $O/r02_c6_bench_numa.json 2> $O/r02_c6_bench_numa.err
echo done
