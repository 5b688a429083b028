#!/bin/bash
# Upper bound of what a faster evaluator role could give: the kernel with the evaluations skipped (TC_EXP_SKIP_EVAL: rows are still
# copied out of TMEM, nothing is evaluated; results are empty), TMA and direct data paths.
for defs in "" "-DTC_EXP_SKIP_EVAL"; do
  SYLDET_TC_DEFS="$defs" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  for env in "SYLDET_TC_DIRECT=0" "SYLDET_TC_DIRECT=1"; do
    env $env timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --no-alt --no-cpu --no-stream --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('defs [$defs] $env: kernel_ms %.3f frac %.3f det %d' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['detections_per_step']))"
  done
done
SYLDET_TC_DEFS="" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1
