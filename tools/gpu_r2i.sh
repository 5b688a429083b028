#!/bin/bash
# Round-2 GPU pass I: resamplers (batched linear, polyphase, CLI formats), wide path after the STFT load fix, full test suite.
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 400 python bench.py --config 4 --hidden 256 --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_c4_h256.json 2> gpurun_out/${TAG}_bench_c4_h256.err; tail -2 gpurun_out/${TAG}_bench_c4_h256.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_c4_h256.json"))
print("H=256 value %.4g ms/step %.2f stft_ms %.2f l0_ms %.2f frac %.3f mma_frac %.3f parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["stft_kernel_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["mma_frac"], d["parity"]))
PY
