#!/bin/bash
# Round-2 GPU pass D: 16 vs 8 spectrum warps (same box), wide-hidden kernel bring-up, full parity tests.
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout -s KILL 120 python tools/tc_check.py 5 2 > gpurun_out/${TAG}_tc_check.log 2>&1; rc=$?; tail -4 gpurun_out/${TAG}_tc_check.log
if [ $rc -ne 0 ]; then echo "tc_check failed rc=$rc"; exit 1; fi
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "wide" 2>&1 | tail -30 | tee gpurun_out/${TAG}_wide_tests.log
bash tools/tc_variants.sh ${TAG} "" "-DTC_NUM_D=8" 2>&1 | tee gpurun_out/${TAG}_variants.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_high_overlap_wide_hidden_config --deselect tests/test_gpu_parity.py::test_wide_kernel_shape_variants 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
