#!/bin/bash
# Quick GPU-box pass while iterating on the tensor kernel: bring-up check, parity tests, device-only bench, role cycle timers.
#   tools/gpu_quick.sh <tag> [pytest -k expression]
TAG=${1:-q}
mkdir -p gpurun_out
timeout -s KILL 90 python tools/tc_check.py 5 2 > gpurun_out/${TAG}_tc_check.log 2>&1; rc=$?; tail -8 gpurun_out/${TAG}_tc_check.log
if [ $rc -ne 0 ] || ! grep -q "off by > 1e-5: 0" gpurun_out/${TAG}_tc_check.log; then echo "tc_check failed (rc=$rc): stopping"; exit 1; fi
if [ -n "$2" ]; then
  timeout -s KILL 600 python -m pytest tests -m gpu -q -x -k "$2" 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
else
  timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
fi
timeout -s KILL 200 python bench.py --no-cpu --no-stream --no-pcm16 --e2e-steps 1 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g  kernel_ms %.3f  frac %.3f  parity %s clocks %s kernel %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["parity"], d["clocks"], d["kernel"]))
PY
SYLDET_TC_TIMING=1 timeout -s KILL 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-stream --no-pcm16 --e2e-steps 1 2> gpurun_out/${TAG}_role_cycles.txt > /dev/null
tail -24 gpurun_out/${TAG}_role_cycles.txt
