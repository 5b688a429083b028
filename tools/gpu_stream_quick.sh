#!/bin/bash
# Quick live-path pass: stream/detector tests only, then the latency driver at 64 x 32 and 1 x 32.
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "stream or syllable_detector or track_detector" 2>&1 | tail -5 | tee gpurun_out/${TAG}_tests.log
B=syllable-detector-swift_b200/syldet_stream_bench
timeout 300 $B -n tests/golden/sample.txt -c 64 -b 32 -s 60 -p 5 | tee gpurun_out/${TAG}_stream_b32.json
timeout 300 $B -n tests/golden/sample.txt -c 1 -b 32 -s 20 -p 0 | tee gpurun_out/${TAG}_stream_c1.json
timeout 300 $B -n tests/golden/sample.txt -c 1024 -b 32 -s 20 -p 0 | tee gpurun_out/${TAG}_stream_c1024.json
