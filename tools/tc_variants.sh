#!/bin/bash
# Build the tensor kernel with different -D switches ON THE GPU BOX and time each (bench.py, device-resident, no CPU leg).
#   tools/tc_variants.sh "<defs A>" "<defs B>" ...     ("" = default build)
mkdir -p gpurun_out
for defs in "$@"; do
  SYLDET_TC_DEFS="$defs" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  timeout 300 python bench.py --kernel tensor --no-cpu --e2e-steps 1 --steps 10 --warmup 3 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('defs [%s]  kernel_ms %.3f  frac %.3f  err %.2e flips %d det %d' % ('$defs', d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['max_abs_err_vs_oracle'], d['parity']['decision_flips'], d['detections_per_step']))"
done
SYLDET_TC_DEFS="" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1
