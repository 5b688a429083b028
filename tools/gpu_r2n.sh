#!/bin/bash
# Round-2 GPU pass N: non-blocking MMA issuer (v23) against the previous build (tools/ab/kernels_tc_v22.cu) on the same box.
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout -s KILL 120 python tools/tc_check.py 5 2 > gpurun_out/${TAG}_tc_check.log 2>&1; rc=$?; tail -4 gpurun_out/${TAG}_tc_check.log
if [ $rc -ne 0 ]; then echo "tc_check failed rc=$rc"; exit 1; fi
bash tools/tc_variants.sh ${TAG} "" 2>&1 | tee gpurun_out/${TAG}_variants.txt
cp syllable-detector-swift_b200/csrc/kernels_tc.cu /tmp/kernels_tc_new.cu
cp tools/ab/kernels_tc_v22.cu syllable-detector-swift_b200/csrc/kernels_tc.cu
bash tools/tc_variants.sh ${TAG}_v22 "" 2>&1 | tee -a gpurun_out/${TAG}_variants.txt
cp /tmp/kernels_tc_new.cu syllable-detector-swift_b200/csrc/kernels_tc.cu
python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -k "tensor or sample or amplitude or spectra or golden or chunk or slices or large or edge" 2>&1 | tail -5 | tee gpurun_out/${TAG}_tests.log
