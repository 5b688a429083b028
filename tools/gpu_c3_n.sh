#!/bin/bash
# Corpus run (config 3) on N GPUs only: tools/gpu_c3_n.sh <tag> <n>
TAG=$1; n=$2
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --config 3 --corpus-hours 1000 > gpurun_out/${TAG}_c3_n${n}.json 2> gpurun_out/${TAG}_c3_n${n}.err
tail -2 gpurun_out/${TAG}_c3_n${n}.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_c3_n${n}.json')); print('value %.4g total_s %.2f gather_s %.2f dev_only %.4g flips %s events %s' % (d['value'], d['seconds_total'], d['gather_seconds'], d['value_device_only'], d['parity']['decision_flips_outside_near_band'], d.get('events_gathered')))"
