"""Where the resident-kernel test spends its time: per-call wall time of submit() on a resident group, alone and next to a launched group."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
sd = importlib.import_module("syllable-detector-swift_b200")
cfg = sd.SyllableDetectorConfig(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "sample.txt")).validate()
nch, nbuf, ticks = 12, 32, 400
x = (0.05 * np.random.default_rng(0).standard_normal((nch, nbuf * ticks))).astype(np.float32)
for both in (False, True):
    ref = sd.StreamGroup(cfg, nch, max_buffer=nbuf) if both else None
    os.environ["SYLDET_STREAM_RESIDENT"] = "1"
    os.environ["SYLDET_STREAM_RESIDENT_IDLE_MS"] = "5"
    g = sd.StreamGroup(cfg, nch, max_buffer=nbuf)
    del os.environ["SYLDET_STREAM_RESIDENT"]
    ta, tb = [], []
    for t in range(ticks):
        buf = x[:, t * nbuf:(t + 1) * nbuf]
        if ref is not None:
            a = time.perf_counter(); ref.submit(buf); ta.append(time.perf_counter() - a)
        a = time.perf_counter(); g.submit(buf); tb.append(time.perf_counter() - a)
    tb = np.array(tb) * 1e6
    print("with launched group beside it" if both else "resident group alone", "resident submit us: p50 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(np.percentile(tb, [50, 90, 99, 100])),
          "launches", g.launch_count, "resident ticks", g.resident_tick_count)
    if ta:
        ta = np.array(ta) * 1e6
        print("   launched submit us: p50 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(np.percentile(ta, [50, 90, 99, 100])))
    del g, ref

# the events of tests/test_gpu_parity.py::test_stream_group_resident_kernel, timed one by one
for rate in (0.0, 48000.0):
    os.environ["SYLDET_STREAM_RESIDENT"] = "1"
    kw = dict(max_buffer=nbuf)
    if rate:
        kw["input_rate"] = rate
    t0 = time.perf_counter(); g = sd.StreamGroup(cfg, nch, **kw); t_create = time.perf_counter() - t0
    del os.environ["SYLDET_STREAM_RESIDENT"]
    slow = []
    for t in range(ticks):
        buf = x[:, t * nbuf:(t + 1) * nbuf]
        a = time.perf_counter(); g.submit(buf); d = time.perf_counter() - a
        if d > 1e-3:
            slow.append((t, round(d * 1e3, 2)))
        if t == 100:
            time.sleep(0.05)
        if t == 200:
            a = time.perf_counter(); g.read_levels(); print("  read_levels ms %.2f" % ((time.perf_counter() - a) * 1e3))
    a = time.perf_counter(); del g; t_del = time.perf_counter() - a
    print("rate", rate, "create ms %.1f, destroy ms %.1f, submits slower than 1 ms (tick, ms):" % (t_create * 1e3, t_del * 1e3), slow[:20], len(slow))
