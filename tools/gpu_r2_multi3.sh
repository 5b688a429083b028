#!/bin/bash
# Session-2 multi-GPU pass (gpurun --gpus 8): the default bench line and the corpus run at N = 8 (and N = 4 when asked for).
TAG=${1:-r02m8}
shift
mkdir -p gpurun_out
run() {  # n, tag, args...
  local n=$1 t=$2; shift 2
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > gpurun_out/${TAG}_${t}_n${n}.json 2> gpurun_out/${TAG}_${t}_n${n}.err
  tail -2 gpurun_out/${TAG}_${t}_n${n}.err | cut -c1-300
}
for n in "$@"; do
  run $n bench --steps 10 --warmup 3
  run $n c3 --config 3 --corpus-hours 1000
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*_n*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    if d.get("scaling") == "strong":
        print(f, "value %.4g total_s %.2f gather_s %.2f dev_only %.4g flips %s" % (d["value"], d["seconds_total"], d["gather_seconds"], d["value_device_only"], d["parity"]["decision_flips_outside_near_band"]))
    else:
        e = d["e2e"]
        print(f, "value %.4g frac %.3f e2e %.4g (%.1f ms, %.1f GB/s per GPU, ceiling %.1f, frac %.2f) e2e_f32 %.4g stream p99 %s" % (d["value"], d["roofline"]["frac"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e["pcie_peak_gbs"], e["pcie_frac"], d.get("e2e_f32", {}).get("value", 0), d.get("stream", {}).get("burst", {}).get("per_buffer", {}).get("p99_us")))
PY
