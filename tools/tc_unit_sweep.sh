#!/bin/bash
# Unit length of the tensor kernel (tiles per unit) A/B on one box: SYLDET_TC_UNIT_TILES forces a value, "" = the planner's choice.
mkdir -p gpurun_out
for envs in "SYLDET_TC_UNIT_TILES=32" "" "SYLDET_TC_UNIT_TILES=38" "SYLDET_TC_UNIT_TILES=20" "SYLDET_TC_UNIT_TILES=32" ""; do
  env $envs timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --no-cpu --no-stream --steps 20 --warmup 3 2> gpurun_out/unit_sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('env [$envs]  kernel_ms %.4f  frac %.4f  tf32 %.3f  err %.2e flips far %d det %d' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('tf32_kernel_ms') or 0, d['parity']['max_abs_err_vs_oracle'], d['parity']['decision_flips_outside_near_band'], d['detections_per_step']))" || tail -5 gpurun_out/unit_sweep.err
done
