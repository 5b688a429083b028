#!/bin/bash
# SIMT fused kernel after a change: parity suite, the sample shape forced onto it, and the FFT-512 shape.
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout -s KILL 300 python bench.py --kernel fused --no-e2e --quick-parity --no-alt --no-cpu --no-stream --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('sample on fused: kernel_ms %.3f frac %.3f err %.2e' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['max_abs_err_vs_oracle']))"
for shape in fft512_hop256_h8 fft256_hop128_h8_minmax; do
  timeout -s KILL 300 python bench.py --shape $shape --no-e2e --no-cpu --no-stream --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$shape: %s kernel_ms %.3f frac %.3f err %.2e flips_far %d' % (d['kernel'][:14], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['max_abs_err_vs_oracle'], d['parity']['decision_flips_outside_near_band']))"
done
