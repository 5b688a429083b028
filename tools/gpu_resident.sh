#!/bin/bash
# Resident tick kernel: its test (bounded: a polling kernel that never leaves must not hold the box), then the latency driver with it.
timeout -s KILL 180 python -m pytest tests -m gpu -q -x -k "resident" 2>&1 | tail -8
B=syllable-detector-swift_b200/syldet_stream_bench
SYLDET_STREAM_RESIDENT=1 SYLDET_STREAM_TIMING=1 timeout -s KILL 120 $B -n tests/golden/sample.txt -c 64 -b 32 -s 20 -p 0 2>&1 | grep -v '^{' | tail -4
SYLDET_STREAM_RESIDENT=1 timeout -s KILL 200 $B -n tests/golden/sample.txt -c 64 -b 32 -s 60 -p 5 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print('burst', d['burst']['per_buffer'], d['burst']['per_buffer_with_new_outputs'], 'rt', d['burst'].get('realtime_factor'))
        if 'paced' in d: print('paced', d['paced'])
        print({k: d[k] for k in d if 'launch' in k or 'resident' in k})
"
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
