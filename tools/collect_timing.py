"""Where the time of one corpus step goes: launch, C collect, Python Events conversion (1 h x 8 ch, sample.txt)."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sd = importlib.import_module("syldet_b200")
synth = importlib.import_module("tools.synth")
pkg = sys.modules[sd.BatchDetector.__module__]
lib = pkg.lib

cfg = sd.SyllableDetectorConfig(os.path.join(ROOT, "tests", "golden", "sample.txt")).validate()
nch, n = 8, 3600 * 44100
dev = torch.device("cuda:0")
x = synth.make_audio_torch(nch, n, dev, seed=5000)
det = sd.BatchDetector(cfg, device=0)
stream = torch.cuda.current_stream(dev)
E = cfg.num_evals(n)
d_out = torch.empty((nch, E, cfg.net_outputs), dtype=torch.float32, device=dev)
for rep in range(6):
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    det.launch_device(x.data_ptr(), nch, n, n, d_outputs_ptr=d_out.data_ptr(), stream=stream.cuda_stream)
    t1 = time.perf_counter()
    ev = C.c_void_p()
    pkg._check(lib.syldet_batch_collect(det._h, 0, C.byref(ev)))
    t2 = time.perf_counter()
    e = sd.Events(ev, cfg.sampling_rate) if hasattr(sd, "Events") else pkg.Events(ev, cfg.sampling_rate)
    t3 = time.perf_counter()
    print("rep %d: launch call %.2f ms, C collect %.2f ms, Events %.2f ms, total %.2f ms (%d events)" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0), len(e)), flush=True)
