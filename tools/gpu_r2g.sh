#!/bin/bash
TAG=${1:-r02g}
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I syllable-detector-swift_b200/csrc tools/mma_rate.cu -o /tmp/mma_rate && timeout -s KILL 120 /tmp/mma_rate --nosw | tee gpurun_out/${TAG}_mma_nosw.txt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -k "stream or resampl" 2>&1 | tail -30 | tee gpurun_out/${TAG}_stream_tests.log
