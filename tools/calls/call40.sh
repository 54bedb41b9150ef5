#!/bin/bash
# GPU call 40 of round 2: the 4K-input property test.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "4k_input" > gpurun_out/r02_c40_pytest_4k.txt 2>&1; tail -15 gpurun_out/r02_c40_pytest_4k.txt
