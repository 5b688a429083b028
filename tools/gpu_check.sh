#!/bin/bash
# One GPU-box pass: parity tests, bench line, launch list, optional full ncu capture of the fused kernel.
#   tools/gpu_check.sh [tag] [--full]
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g audio-s/s  frames/s %.4g  kernel_ms %.3f  roofline.frac %.3f  e2e %.4g  cpu %.4g (%s cores)  parity %s clocks %s" % (
  d["value"], d["frames_per_s"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d.get("cpu_baseline",{}).get("value",0), d.get("cpu_baseline",{}).get("cores"), d["parity"], d["clocks"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread --clock-control none -k regex:detect_kernel -s 3 -c 2 --csv --log-file gpurun_out/${TAG}_ncu_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu --hours 0.25 > /dev/null 2>&1
grep -E "detect_kernel" gpurun_out/${TAG}_ncu_metrics.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | head -14
if [ "$2" == "--full" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-cpu --hours 0.25 > gpurun_out/${TAG}_ncu_full.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_full.log
fi
