#!/bin/bash
# Time the default tensor kernel under different run-time switches ON THE GPU BOX (no rebuild): for each "VAR=val VAR=val" set,
# production timing (bench.py, device-resident, no CPU leg) and the role cycle timers.
#   tools/tc_env_sweep.sh <tag> "<env A>" "<env B>" ...     ("" = defaults)
TAG=$1; shift
mkdir -p gpurun_out
n=0
for envs in "$@"; do
  n=$((n+1))
  env $envs timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --steps 10 --warmup 3 2> gpurun_out/${TAG}_e${n}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('env [$envs]  kernel_ms %.3f  frac %.3f  tf32 %.3f  err %.2e flips %d (far %d) det %d fallbacks %d clocks %s' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['tf32_kernel_ms'] or 0, d['parity']['max_abs_err_vs_oracle'], d['parity']['decision_flips'], d['parity']['decision_flips_outside_near_band'], d['detections_per_step'], d['range_fallbacks'], d['clocks']))" || tail -5 gpurun_out/${TAG}_e${n}.err
  env $envs SYLDET_TC_TIMING=1 timeout -s KILL 160 python bench.py --kernel tensor --no-e2e --quick-parity --no-alt --steps 2 --warmup 3 2>&1 >/dev/null | grep -A23 "1038 tiles" | tail -24 > gpurun_out/${TAG}_e${n}_cycles.txt
  python - <<PY
rows=[l.split() for l in open("gpurun_out/${TAG}_e${n}_cycles.txt")]
print("   ", rows[0][-14:-6] if rows else "")
print("   ", " ".join("%s=%s" % (r[0], r[1]) for r in rows[1:] if len(r)==2))
PY
done
