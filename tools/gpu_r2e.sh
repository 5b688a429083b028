#!/bin/bash
# Round-2 GPU pass E: peak microbenchmarks, config-4 bench (H = 256 / 1024), ncu captures of the tensor kernel and the wide kernels.
TAG=${1:-r02e}
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I syllable-detector-swift_b200/csrc tools/mma_rate.cu -o /tmp/mma_rate && timeout -s KILL 120 /tmp/mma_rate --peak | tee gpurun_out/${TAG}_tensor_peaks.json
nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ffma_peak.cu -o /tmp/ffma_peak && timeout -s KILL 120 /tmp/ffma_peak | tee gpurun_out/${TAG}_ffma_peak.json
cp gpurun_out/${TAG}_tensor_peaks.json profiles/r02_tensor_peaks.json
timeout -s KILL 400 python bench.py --config 4 --hidden 256 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c4_h256.json 2> gpurun_out/${TAG}_bench_c4_h256.err; tail -2 gpurun_out/${TAG}_bench_c4_h256.err; cut -c1-3000 gpurun_out/${TAG}_bench_c4_h256.json
timeout -s KILL 400 python bench.py --config 4 --hidden 1024 --steps 3 --warmup 3 --no-cpu --wide-seconds 10 > gpurun_out/${TAG}_bench_c4_h1024.json 2> gpurun_out/${TAG}_bench_c4_h1024.err; tail -2 gpurun_out/${TAG}_bench_c4_h1024.err; cut -c1-2500 gpurun_out/${TAG}_bench_c4_h1024.json
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"wide_l0_kernel|stft_planes_kernel" -s 4 -c 2 -f -o gpurun_out/${TAG}_wide_prof python bench.py --config 4 --hidden 256 --steps 1 --warmup 3 --no-cpu --wide-seconds 4 --channels 8 > gpurun_out/${TAG}_wide_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_wide_ncu.log
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_tc_prof python bench.py --kernel tensor --no-e2e --quick-parity --no-alt --steps 1 --warmup 3 --hours 0.25 > gpurun_out/${TAG}_tc_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_tc_ncu.log
timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g kernel_ms %.3f frac %.3f variants %s e2e %.4g (pcie_frac %.2f) e2e_f32 %.4g cpu %s parity %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"].get("variants"), d["e2e"]["value"], d["e2e"]["pcie_frac"], d.get("e2e_f32",{}).get("value",0), d.get("cpu_baseline"), {k:v for k,v in d["parity"].items() if k!="near_threshold_frames"}))
print("stream", json.dumps(d.get("stream"))[:1500])
PY
