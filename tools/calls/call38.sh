#!/bin/bash
# GPU call 38 of round 2: the same arithmetic on three kinds of input frames: how much of the power cap is switching activity.
set -u
mkdir -p gpurun_out
for k in noise edges real noise edges real; do
  timeout 300 python bench.py --frames $k --no-cpu >> gpurun_out/r02_c38_bench_frames.jsonl 2>> gpurun_out/r02_c38_bench_frames.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_c38_bench_frames.jsonl'):
    d=json.loads(l); print(d['config']['frames'], round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), d['roofline']['ms_per_frame'])
PY
