#!/bin/bash
# Round-2 session-3 pass: full parity suite, default bench line (both arms), stream sweep inside it.
TAG=${1:-r02v26}
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 500 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout -s KILL 700 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g ms %.3f frac %.3f e2e %.4g cpu %.4g launches %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"]))
print("stream", json.dumps(d.get("stream"))[:1500])
r=json.load(open("gpurun_out/${TAG}_bench_reference.json")); print("reference", r["value"], r["cpu_baseline"])
PY
