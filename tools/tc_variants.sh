#!/bin/bash
# Build the tensor kernel with different -D switches ON THE GPU BOX; for each: production timing (bench.py, device-resident,
# no CPU leg), role cycle timers (cycles per tile, effective SM clock) and optionally one full ncu capture.
#   [NCU=1] tools/tc_variants.sh <tag> "<defs A>" "<defs B>" ...     ("" = default build)
TAG=$1; shift
mkdir -p gpurun_out
n=0
for defs in "$@"; do
  n=$((n+1))
  SYLDET_TC_DEFS="$defs" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --steps 10 --warmup 3 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('defs [$defs]  kernel_ms %.3f  frac %.3f  tf32 %.3f  err %.2e flips %d (far %d) det %d fallbacks %d clocks %s' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['tf32_kernel_ms'] or 0, d['parity']['max_abs_err_vs_oracle'], d['parity']['decision_flips'], d['parity']['decision_flips_outside_near_band'], d['detections_per_step'], d['range_fallbacks'], d['clocks']))"
  SYLDET_TC_TIMING=1 timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --no-alt --steps 2 --warmup 3 2>&1 >/dev/null | grep -A23 "1038 tiles" | tail -24 > gpurun_out/${TAG}_v${n}_cycles.txt
  python - <<PY
rows=[l.split() for l in open("gpurun_out/${TAG}_v${n}_cycles.txt")]
print("   ", rows[0][-14:-6] if rows else "")
print("   ", " ".join("%s=%s" % (r[0], r[1]) for r in rows[1:] if len(r)==2))
PY
  if [ -n "$NCU" ]; then
    timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_v${n}_prof python bench.py --kernel tensor --no-e2e --quick-parity --no-alt --steps 1 --warmup 3 --hours 0.25 > gpurun_out/${TAG}_v${n}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_v${n}_ncu.log
  fi
done
SYLDET_TC_DEFS="" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1
