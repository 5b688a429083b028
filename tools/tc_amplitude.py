"""Amplitude sweep for the tensor-core kernel: the fp16 correction pass is exact to fp32 level only while samples are fp16
normals (6e-5 <= |x| <= 65504). Prints max |tensor - oracle| per scale; run again with SYLDET_TC_TF32_CORR=1 to compare."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sd = importlib.import_module("syldet_b200")
synth = importlib.import_module("tools.synth")
from oracle import Oracle

path = os.path.join(ROOT, "tests", "golden", "sample.txt")
cfg = sd.SyllableDetectorConfig(path).validate()
orc = Oracle(path)
x0 = synth.make_audio(2, 44100 * 4, seed=5)
print("corrections:", "tf32" if os.environ.get("SYLDET_TC_TF32_CORR") else "fp16", " rms of the unscaled audio %.2e" % float(np.sqrt(np.mean(x0 ** 2))))
for e in (14, 10, 0, -6, -10, -14, -18, -22):
    x = (x0 * np.float32(2.0 ** e)).astype(np.float32)
    ev, out = sd.BatchDetector(cfg, kernel=sd.KERNEL_TENSOR).run(x, want_outputs=True)
    ref = np.stack([orc.run(x[ch])[0] for ch in range(x.shape[0])])
    d = np.abs(out - ref)
    print("scale 2^%-4d max err %.3e  mean %.3e  nan %d  (> 1e-5: %d of %d)" % (e, np.nanmax(d), np.nanmean(d), int(np.isnan(out).sum()), int((d > 1e-5).sum()), d.size), flush=True)
