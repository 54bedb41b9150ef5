#!/bin/bash
# GPU call 10 of round 2: long determinism soak of the shipped build (chained kernel, cooperative launch, balanced split).
set -u
mkdir -p gpurun_out
timeout 900 python tools/race_hunt.py 6000 > gpurun_out/r02_c10_race_hunt_6000.txt 2>&1
echo done
