#!/bin/bash
# Build the wide path with different -D switches ON THE GPU BOX and time config 4 (H = 256): STFT-planes kernel and contraction.
#   tools/wide_variants.sh <tag> "<defs A>" "<defs B>" ...     ("" = default build)
TAG=$1; shift
mkdir -p gpurun_out
for defs in "$@"; do
  SYLDET_WIDE_DEFS="$defs" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  timeout -s KILL 300 python bench.py --config 4 --hidden 256 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('defs [$defs] ms/step %.3f stft %.3f contraction %.3f value %.5g err %.2e flips %d' % (d['ms_per_step'], d['roofline']['stft_kernel_ms'], d['roofline']['kernel_ms'], d['value'], d['parity']['max_abs_err_vs_oracle'], d['parity'].get('decision_flips', -1)))"
done
SYLDET_WIDE_DEFS="" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1
