#!/bin/bash
# Is the splitters' load latency DRAM latency? Role cycle timers with a recording that fits in L2 (hours 0.01: 51 MB) against
# the usual one, per build switch.   tools/tc_l2_probe.sh <tag> "<defs A>" ...
TAG=$1; shift
for defs in "$@"; do
  SYLDET_TC_DEFS="$defs" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  for h in 0.01 1.0; do
    for pf in 0 2; do
      echo "== defs [$defs] hours $h pf $pf"
      SYLDET_TC_PF=$pf timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --no-alt --no-cpu --no-stream --steps 10 --warmup 3 --hours $h 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('kernel_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['roofline']['frac']))"
      SYLDET_TC_PF=$pf SYLDET_TC_TIMING=1 timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --no-alt --no-cpu --no-stream --steps 2 --warmup 3 --hours $h 2>&1 >/dev/null | grep -A23 "tiles per CTA" | tail -24 | awk '{printf "%s=%s ", $1, $2} END {print ""}' | cut -c1-700
    done
  done
done
SYLDET_TC_DEFS="" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1
