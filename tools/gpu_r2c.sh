#!/bin/bash
# Round-2 GPU pass C: 16 spectrum warps vs 8 (same box), tensor-kernel parity tests.
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout -s KILL 120 python tools/tc_check.py 5 2 > gpurun_out/${TAG}_tc_check.log 2>&1; rc=$?; tail -6 gpurun_out/${TAG}_tc_check.log
if [ $rc -ne 0 ]; then echo "tc_check failed rc=$rc"; exit 1; fi
bash tools/tc_variants.sh ${TAG} "" "-DTC_NUM_D=8" 2>&1 | tee gpurun_out/${TAG}_variants.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x -k "tensor or sample or amplitude or spectra or golden or chunk or slices or large" 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
