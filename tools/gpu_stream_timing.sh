#!/bin/bash
# Live-path diagnostics: device cycle stamps per phase + host microseconds per step (SYLDET_STREAM_TIMING=1).
timeout 600 python -m pytest tests -m gpu -q -x -k "stream or syllable_detector or track_detector or bit_exact or generated" 2>&1 | tail -5
B=syllable-detector-swift_b200/syldet_stream_bench
for c in 1 64; do SYLDET_STREAM_TIMING=1 timeout 300 $B -n tests/golden/sample.txt -c $c -b 32 -s 20 -p 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['channels'], d['burst']['per_buffer_with_new_outputs'])
    else: print(l)
"; done
