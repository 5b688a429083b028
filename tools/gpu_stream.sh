#!/bin/bash
# Live-path GPU pass: parity tests, then the per-buffer latency driver at 32/128/256-frame buffers (BASELINE config 5).
#   tools/gpu_stream.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
B=syllable-detector-swift_b200/syldet_stream_bench
for nb in 32 128 256; do
  timeout 300 $B -n tests/golden/sample.txt -c 64 -b $nb -s 60 -p 5 > gpurun_out/${TAG}_stream_b${nb}.json 2> gpurun_out/${TAG}_stream_b${nb}.err
  cat gpurun_out/${TAG}_stream_b${nb}.json
done
timeout 300 $B -n tests/golden/sample.txt -c 1 -b 32 -s 20 -p 0 | tee gpurun_out/${TAG}_stream_c1.json
timeout 300 $B -n tests/golden/sample.txt -c 1024 -b 32 -s 20 -p 0 | tee gpurun_out/${TAG}_stream_c1024.json
