#!/bin/bash
# GPU call 29 of round 2: first conv alone after slimming the MMA issuer loop; parity check script; in-pipeline A/B vs im2col.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02_c29_conv0_alone.txt
: > $O
for i in 1 2; do
  timeout 120 python tools/time_conv0.py >> $O 2>&1
  timeout 120 python tools/time_conv0.py 1920x1080x2 128 >> $O 2>&1
done
cat $O
timeout 600 python tools/check_conv0_rows.py > gpurun_out/r02_c29_conv0_rows.txt 2> gpurun_out/r02_c29_conv0_rows.err; grep -c '"oracle_bad_frac": 0.0' gpurun_out/r02_c29_conv0_rows.txt; tail -7 gpurun_out/r02_c29_conv0_rows.txt | cut -c1-120
