#!/bin/bash
# Round-2 GPU pass K: end-to-end tail (per-slice row building), second-shape benches, launch list, full bench.
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -k "slices or large or overflow or debounce or cli or pcm16 or polyphase or amplitude" 2>&1 | tail -8 | tee gpurun_out/${TAG}_tests.log
for shape in fft512_hop256_h8 fft256_hop128_h8_minmax; do
  timeout -s KILL 300 python bench.py --shape $shape --quick-parity --steps 5 --warmup 3 > gpurun_out/${TAG}_shape_${shape}.json 2> gpurun_out/${TAG}_shape_${shape}.err; tail -1 gpurun_out/${TAG}_shape_${shape}.err | cut -c1-300; cut -c1-900 gpurun_out/${TAG}_shape_${shape}.json
done
SYLDET_E2E_TIMING=1 timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; grep "syldet e2e" gpurun_out/${TAG}_bench.err | tail -6
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g kernel_ms %.3f frac %.3f variants %s e2e %.4g (%.1f ms, pcie_frac %.2f) e2e_f32 %.4g cpu %.4g single %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], {k[:11]: round(v["frac"],3) for k,v in d["roofline"].get("variants",{}).items()}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pcie_frac"], d.get("e2e_f32",{}).get("value",0), d["cpu_baseline"]["value"], d["cpu_baseline"]["single_core"]["value"]))
PY
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-stream --e2e-steps 1 --no-f32-e2e --quick-parity > /dev/null 2>&1
grep -c . gpurun_out/${TAG}_launches.csv
