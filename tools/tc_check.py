"""Bring-up check for the tensor-core kernel: compare against the SIMT fused kernel and the oracle on a small input."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sd = importlib.import_module("syldet_b200")
synth = importlib.import_module("tools.synth")
from oracle import Oracle

cfg = sd.SyllableDetectorConfig(os.path.join(ROOT, "tests", "golden", "sample.txt")).validate()
orc = Oracle(os.path.join(ROOT, "tests", "golden", "sample.txt"))
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
nch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = synth.make_audio(nch, int(44100 * secs), seed=3)
print("available kernels", sd.BatchDetector.available_kernels(cfg), flush=True)
ev_f, out_f = sd.BatchDetector(cfg, kernel=sd.KERNEL_FUSED).run(x, want_outputs=True)
print("fused done", out_f.shape, flush=True)
ev_t, out_t = sd.BatchDetector(cfg, kernel=sd.KERNEL_TENSOR).run(x, want_outputs=True)
print("tensor done", out_t.shape, flush=True)
ref = np.stack([orc.run(x[ch])[0] for ch in range(nch)])
d_tf = np.abs(out_t - out_f)
d_to = np.abs(out_t - ref)
print("tensor vs fused  max %.3e  mean %.3e   nan %d" % (np.nanmax(d_tf), np.nanmean(d_tf), np.isnan(out_t).sum()))
print("tensor vs oracle max %.3e  fused vs oracle max %.3e" % (np.nanmax(d_to), np.nanmax(np.abs(out_f - ref))))
bad = np.argwhere(~(d_to <= 1e-5))
print("evaluations off by > 1e-5:", len(bad), bad[:10].tolist())
print("first outputs tensor", out_t[0, :5, 0], "oracle", ref[0, :5, 0])
print("events tensor/fused", len(ev_t), len(ev_f), "same samples:", np.array_equal(ev_t.sample, ev_f.sample) and np.array_equal(ev_t.channel, ev_f.channel))
