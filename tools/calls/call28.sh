#!/bin/bash
# GPU call 28 of round 2: ncu --set full of the row-streaming first conv alone (one launch).
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv0_rows -s 3 -c 1 -f -o gpurun_out/r02_c28_conv0_rows \
  python tools/time_conv0.py > gpurun_out/r02_c28_ncu.log 2>&1
tail -5 gpurun_out/r02_c28_ncu.log
ls -la gpurun_out/*.ncu-rep
