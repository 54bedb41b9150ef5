#!/bin/bash
# GPU call 23 of round 2: the row-streaming first conv (conv0_rows.cu) against the im2col one: parity, then device time.
set -u
mkdir -p gpurun_out
timeout 600 python tools/check_conv0_rows.py > gpurun_out/r02_c23_conv0_rows.txt 2> gpurun_out/r02_c23_conv0_rows.err
echo rc=$?
tail -20 gpurun_out/r02_c23_conv0_rows.txt
tail -5 gpurun_out/r02_c23_conv0_rows.err
