#!/bin/bash
# Round-2 GPU pass B: all-fp16 band DFT. Bring-up check, variants (barrier wait policy), parity tests.
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout -s KILL 120 python tools/tc_check.py 5 2 > gpurun_out/${TAG}_tc_check.log 2>&1; rc=$?; tail -6 gpurun_out/${TAG}_tc_check.log
if [ $rc -ne 0 ]; then echo "tc_check failed rc=$rc"; exit 1; fi
bash tools/tc_variants.sh ${TAG} "" "-DSYLDET_MBAR_SLEEP=100" "-DSYLDET_MBAR_HINT=2000" 2>&1 | tee gpurun_out/${TAG}_variants.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/${TAG}_tests.log
