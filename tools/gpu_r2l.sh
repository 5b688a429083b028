#!/bin/bash
# Round-2 GPU pass L: full parity suite, default bench (all legs), corpus run, launch list after the ingest / collect changes.
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${TAG}_tests.log
SYLDET_E2E_TIMING=1 timeout -s KILL 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; grep "syldet e2e" gpurun_out/${TAG}_bench.err | tail -3
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.4g kernel_ms %.3f frac %.3f variants %s e2e %.4g (%.1f ms, pcie_frac %.2f) e2e_f32 %.4g cpu %.4g single %.4g stream p99 %s" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], {k[:11]: round(v["frac"],3) for k,v in d["roofline"].get("variants",{}).items()}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pcie_frac"], d.get("e2e_f32",{}).get("value",0), d["cpu_baseline"]["value"], d["cpu_baseline"]["single_core"]["value"], d["stream"]["burst"]["per_buffer"]["p99_us"]))
print({k: (v.get("burst", {}).get("per_buffer", {}).get("p99_us"), v.get("paced", {}).get("per_buffer", {}).get("p99_us"), v.get("burst", {}).get("realtime_factor")) for k, v in d["stream"].get("sweep", {}).items()})
PY
SYLDET_E2E_TIMING=1 timeout -s KILL 600 python bench.py --config 3 --corpus-hours 1000 > gpurun_out/${TAG}_bench_c3_n1.json 2> gpurun_out/${TAG}_bench_c3_n1.err; grep "syldet e2e" gpurun_out/${TAG}_bench_c3_n1.err | tail -2; cut -c1-200 gpurun_out/${TAG}_bench_c3_n1.json; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_c3_n1.json')); print('c3 value %.4g total_s %.2f gather_s %.2f' % (d['value'], d['seconds_total'], d['gather_seconds']))"
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-stream --e2e-steps 1 --no-f32-e2e --quick-parity > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/${TAG}_launches.csv')))
hdr=None
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        name=r[hdr.index('Kernel Name')][:60]; v=float(r[hdr.index('Metric Value')].replace(',','')); u=r[hdr.index('Metric Unit')]
        v = v/1e3 if u.startswith('us') else v/1e6 if u.startswith('ns') else v
        agg[name][0]+=1; agg[name][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:6]: print('%4d %10.3f ms  %s'%(n,t,k))
PY
