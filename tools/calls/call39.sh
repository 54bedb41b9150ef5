#!/bin/bash
# GPU call 39 of round 2: ncu --set full of the shipped first conv (two rows per step), alone and inside the bench command.
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv0_rows -s 3 -c 1 -f -o gpurun_out/r02_c39_conv0_rows \
  python tools/time_conv0.py > gpurun_out/r02_c39_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:conv0_rows -s 6 -c 1 -f -o gpurun_out/r02_c39_conv0_rows_bench \
  python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02_c39_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
