#!/bin/bash
# Live tick after a kernel change: stream/detector/cli tests, per-phase cycle stamps, then the latency driver at 64 x 32 (burst + paced).
timeout 700 python -m pytest tests -m gpu -q -x -k "stream or detector or cli" 2>&1 | tail -4
B=syllable-detector-swift_b200/syldet_stream_bench
SYLDET_STREAM_TIMING=1 timeout 300 $B -n tests/golden/sample.txt -c 64 -b 32 -s 20 -p 0 2>&1 | grep -v '^{' | tail -4
timeout 300 $B -n tests/golden/sample.txt -c 64 -b 32 -s 60 -p 5 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print('burst', d['burst']['per_buffer'], d['burst']['per_buffer_with_new_outputs'], 'rt', d['burst'].get('realtime_factor'))
        if 'paced' in d: print('paced', d['paced'])
"
