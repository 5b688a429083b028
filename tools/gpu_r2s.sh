#!/bin/bash
# Round-2 session-2 GPU pass: parity suite, config 4 (wide path) and the second shapes after the swizzled FFT buffers / packed adds.
TAG=${1:-r02s2}
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 400 python bench.py --config 4 --hidden 256 --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c4_h256.json 2> gpurun_out/${TAG}_bench_c4_h256.err; tail -2 gpurun_out/${TAG}_bench_c4_h256.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_c4_h256.json"))
print("c4 h256: value %.5g ms/step %.3f frac %.3f phases %s parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("phases_ms") or d.get("phases_ms"), {k: d["parity"][k] for k in ("max_abs_err_vs_oracle","decision_flips") if k in d["parity"]}))
PY
for shape in fft512_hop256_h8 fft256_hop128_h8_minmax; do
  timeout -s KILL 300 python bench.py --shape $shape --no-e2e --no-cpu --no-stream --steps 10 --warmup 3 > gpurun_out/${TAG}_shape_$shape.json 2> gpurun_out/${TAG}_shape_$shape.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_shape_$shape.json"))
print("$shape: %s value %.4g kernel_ms %.3f frac %.3f err %.2e flips_far %d" % (d["kernel"][:28], d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["parity"]["max_abs_err_vs_oracle"], d["parity"]["decision_flips_outside_near_band"]))
PY
done
