#!/bin/bash
# Round-2 GPU pass A: bring-up of the pair-scheme tensor kernel, parity tests, device-only bench, role timers, A/B against the
# round-1 kernel source (tools/ab/kernels_tc_v12.cu) on the same box, then the full bench line.
#   tools/gpu_r2a.sh <tag>
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt; lscpu | grep -E "Model name|NUMA|Socket" >> gpurun_out/${TAG}_smi.txt
timeout -s KILL 120 python tools/tc_check.py 5 2 > gpurun_out/${TAG}_tc_check.log 2>&1; rc=$?; tail -8 gpurun_out/${TAG}_tc_check.log
if [ $rc -ne 0 ] || ! grep -q "off by > 1e-5: 0" gpurun_out/${TAG}_tc_check.log; then echo "tc_check failed (rc=$rc)"; fi
if [ $rc -eq 0 ]; then
  timeout -s KILL 200 python bench.py --no-e2e --quick-parity --steps 10 --warmup 3 > gpurun_out/${TAG}_dev.json 2> gpurun_out/${TAG}_dev.err; cat gpurun_out/${TAG}_dev.json | cut -c1-1500
  SYLDET_TC_TIMING=1 timeout -s KILL 200 python bench.py --no-e2e --quick-parity --no-alt --steps 1 --warmup 3 2> gpurun_out/${TAG}_role_cycles.txt > /dev/null
  grep -A22 "tc timing" gpurun_out/${TAG}_role_cycles.txt | tail -23
fi
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/${TAG}_tests.log
# A/B: the round-1 kernel on this box
cp syllable-detector-swift_b200/csrc/kernels_tc.cu /tmp/kernels_tc_new.cu
cp tools/ab/kernels_tc_v12.cu syllable-detector-swift_b200/csrc/kernels_tc.cu
python syllable-detector-swift_b200/build.py > /dev/null 2>&1 && timeout -s KILL 200 python bench.py --no-e2e --quick-parity --steps 10 --warmup 3 > gpurun_out/${TAG}_dev_v12.json 2> gpurun_out/${TAG}_dev_v12.err
cut -c1-600 gpurun_out/${TAG}_dev_v12.json
cp /tmp/kernels_tc_new.cu syllable-detector-swift_b200/csrc/kernels_tc.cu
python syllable-detector-swift_b200/build.py > /dev/null 2>&1
if [ $rc -eq 0 ]; then
  timeout -s KILL 200 python bench.py --no-e2e --quick-parity --steps 10 --warmup 3 > gpurun_out/${TAG}_dev2.json 2> gpurun_out/${TAG}_dev2.err; cut -c1-600 gpurun_out/${TAG}_dev2.json
  timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
  cut -c1-6000 gpurun_out/${TAG}_bench.json
fi
