#!/bin/bash
# Build the SIMT fused kernel with different -D switches ON THE GPU BOX; time the sample shape forced onto it and the FFT-512 shape.
#   tools/fused_variants.sh "<defs A>" "<defs B>" ...
for defs in "$@"; do
  SYLDET_FUSED_DEFS="$defs" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  timeout -s KILL 300 python bench.py --kernel fused --no-e2e --quick-parity --no-alt --no-cpu --no-stream --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('defs [$defs] sample on fused: kernel_ms %.3f frac %.3f err %.2e' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['max_abs_err_vs_oracle']))"
  timeout -s KILL 300 python bench.py --shape fft512_hop256_h8 --no-e2e --no-cpu --no-stream --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('defs [$defs] fft512: kernel_ms %.3f frac %.3f err %.2e' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['max_abs_err_vs_oracle']))"
done
SYLDET_FUSED_DEFS="" python syllable-detector-swift_b200/build.py --force > /dev/null 2>&1
