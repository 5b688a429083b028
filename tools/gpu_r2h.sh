#!/bin/bash
# Round-2 GPU pass H: wide kernel with the 8-warp templated epilogue; stream/resampler tests; corpus run with the radix event sort.
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q -x -k "wide or stream or resampl or overflow or debounce or slices" 2>&1 | tail -30 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 400 python bench.py --config 4 --hidden 256 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c4_h256.json 2> gpurun_out/${TAG}_bench_c4_h256.err; tail -2 gpurun_out/${TAG}_bench_c4_h256.err
timeout -s KILL 400 python bench.py --config 4 --hidden 1024 --steps 3 --warmup 3 --no-cpu --wide-seconds 10 > gpurun_out/${TAG}_bench_c4_h1024.json 2> gpurun_out/${TAG}_bench_c4_h1024.err; tail -2 gpurun_out/${TAG}_bench_c4_h1024.err
python - <<PY
import json
for h in (256, 1024):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_c4_h%d.json" % h))
        print("H=%d value %.4g ms/step %.2f stft_ms %.2f l0_ms %.2f frac %.3f mma_frac %.3f e2e %.4g parity %s cpu %s" % (h, d["value"], d["ms_per_step"], d["roofline"]["stft_kernel_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["mma_frac"], d["e2e"]["value"], d["parity"], d.get("cpu_baseline",{}).get("value")))
    except Exception as e:
        print("H=%d failed: %r" % (h, e))
PY
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"wide_l0_kernel|stft_planes_fast_kernel" -s 4 -c 2 -f -o gpurun_out/${TAG}_wide_prof python bench.py --config 4 --hidden 256 --steps 1 --warmup 3 --no-cpu --wide-seconds 4 --channels 8 > gpurun_out/${TAG}_wide_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_wide_ncu.log
timeout -s KILL 600 python bench.py --config 3 --corpus-hours 1000 > gpurun_out/${TAG}_bench_c3_n1.json 2> gpurun_out/${TAG}_bench_c3_n1.err; tail -2 gpurun_out/${TAG}_bench_c3_n1.err; cut -c1-900 gpurun_out/${TAG}_bench_c3_n1.json
