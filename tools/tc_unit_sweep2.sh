#!/bin/bash
# Longer units: is the per-unit cost (pipeline fill, warm-up columns) worth removing?
mkdir -p gpurun_out
for envs in "SYLDET_TC_UNIT_TILES=74" "SYLDET_TC_UNIT_TILES=518" "SYLDET_TC_UNIT_TILES=260" "SYLDET_TC_UNIT_TILES=130" "SYLDET_TC_UNIT_TILES=74" "SYLDET_TC_UNIT_TILES=518"; do
  env $envs timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --no-cpu --no-stream --no-alt --steps 20 --warmup 3 2> gpurun_out/unit_sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('env [$envs]  kernel_ms %.4f  frac %.4f  err %.2e flips far %d det %d' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['max_abs_err_vs_oracle'], d['parity']['decision_flips_outside_near_band'], d['detections_per_step']))" || tail -5 gpurun_out/unit_sweep.err
done
