#!/bin/bash
# Round-2 multi-GPU pass (gpurun --gpus 8): the default bench at N = 8 and 4 (end-to-end scaling, bare-copy ceiling with all ranks
# copying at once, NUMA placement) and the corpus run (config 3) at N = 8, 4, 2 (strong scaling, per-shard parity, event gather).
TAG=${1:-r02m}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1; lscpu | grep -E "Model name|NUMA|Socket|^CPU\(s\)" >> gpurun_out/${TAG}_topo.txt; free -g >> gpurun_out/${TAG}_topo.txt
for f in /sys/bus/pci/devices/*/numa_node; do :; done
run() {  # n, tag, args...
  local n=$1 t=$2; shift 2
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > gpurun_out/${TAG}_${t}_n${n}.json 2> gpurun_out/${TAG}_${t}_n${n}.err
  tail -2 gpurun_out/${TAG}_${t}_n${n}.err | cut -c1-300
}
run 8 c3 --config 3 --corpus-hours 1000
run 8 bench --steps 10 --warmup 3
run 4 bench --steps 10 --warmup 3
run 4 c3 --config 3 --corpus-hours 1000
run 2 c3 --config 3 --corpus-hours 1000
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*_n*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    if d.get("scaling") == "strong":
        print(f, "value %.4g total_s %.2f gather_s %.2f dev_only %.4g parity %s" % (d["value"], d["seconds_total"], d["gather_seconds"], d["value_device_only"], d["parity"]))
    else:
        e = d["e2e"]
        print(f, "value %.4g frac %.3f e2e %.4g (%.1f ms, %.1f GB/s per GPU, ceiling %.1f, frac %.2f, numa %s) e2e_f32 %.4g parity %s" % (d["value"], d["roofline"]["frac"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e["pcie_peak_gbs"], e["pcie_frac"], e["numa"], d.get("e2e_f32", {}).get("value", 0), {k: v for k, v in d["parity"].items() if k != "near_threshold_frames"}))
PY
