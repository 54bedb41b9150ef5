#!/bin/bash
# GPU call 35 of round 2: row-streaming first conv alone: all work stages compiled out (7: no gather / arithmetic / stores; 15: no MMAs
# either) = the cost of the hand-shakes alone; waits without watchdog clock reads (16), sleeping waits (32), both (48).
set -u
mkdir -p gpurun_out
L=$PWD/reve_b200
O=gpurun_out/r02_c35_conv0_handshake_floor.txt
: > $O
for i in 1 2; do
for v in libreve_cuda libreve_cuda_xp7 libreve_cuda_xp15 libreve_cuda_xp16 libreve_cuda_xp32 libreve_cuda_xp48; do
  REVE_LIB=$L/$v.so timeout 120 python tools/time_conv0.py >> $O 2>&1
done
done
cut -c1-170 $O
