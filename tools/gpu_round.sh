#!/bin/bash
# Full GPU-box pass for a round: parity tests, both bench arms, role cycle timers, ncu launch list, one full ncu capture.
#   tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
timeout -s KILL 500 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout -s KILL 700 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | cut -c1-3000
SYLDET_TC_TIMING=1 timeout -s KILL 240 python bench.py --steps 3 --warmup 3 --no-cpu --no-stream --e2e-steps 1 2> gpurun_out/${TAG}_role_cycles.txt > /dev/null
tail -30 gpurun_out/${TAG}_role_cycles.txt
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-stream --e2e-steps 1 > /dev/null 2>&1
grep -c . gpurun_out/${TAG}_launches.csv
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:detect_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-stream --e2e-steps 1 --hours 0.25 > gpurun_out/${TAG}_ncu_full.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_full.log
