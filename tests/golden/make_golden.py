"""Generates tests/golden/golden_cases.npz from the CPU oracle (oracle/oracle.c) in THIS container.

The reference ships no golden vectors for the hot path and cannot be built on Linux (Swift + Accelerate), so these
fixtures pin the oracle itself (regression + transport to the GPU box), not Accelerate: "parity unpinned".
    python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Oracle  # noqa: E402

cw = importlib.import_module("syllable-detector-swift_b200.config_writer")
synth = importlib.import_module("tools.synth")


def cases():
    sample = open(os.path.join(ROOT, "tests", "golden", "sample.txt")).read()
    yield "sample", sample, synth.make_audio(1, 44100 * 2, seed=11)[0]
    rng = np.random.default_rng(5)
    n = 30000
    t = np.arange(n)
    x = (0.02 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 3000 * t / 44100) * (np.sin(2 * np.pi * 3 * t / 44100) > 0)).astype(np.float32)
    yield "gap_db_2out", cw.random_config(seed=3, fft_len=512, win_len=400, overlap=-20, freq_range=(1000, 8000), time_range=6,
                                          hidden=(8, 3), outputs=2, scaling="db", input_funcs=("normalize", "mapstd"),
                                          output_funcs=("mapstd", "mapminmax"), transfer="LogSig", out_transfer="SatLin",
                                          threshold=[0.5, 0.4]), x
    yield "log_std_128", cw.random_config(seed=4, fft_len=128, win_len=128, overlap=96, freq_range=(500, 9000), time_range=4,
                                          hidden=(5,), outputs=1, scaling="log", input_funcs=("normalizestd",),
                                          output_funcs=(), transfer="SatLin", out_transfer="PureLin", threshold=0.1), x[:20000]
    yield "wide_64", cw.random_config(seed=6, fft_len=64, win_len=50, overlap=10, freq_range=(0, 22050), time_range=3,
                                      hidden=(40, 12), outputs=3, scaling="linear", input_funcs=("mapminmax", "l2normalize"),
                                      output_funcs=("mapminmax",), transfer="TanSig", out_transfer="LogSig",
                                      threshold=[0.5, 0.5, 0.5]), x[:12000]


def retarget_thresholds(text, audio, q=0.7):
    """Replace the thresholds by the q-quantile of each output on this audio, so both sides of the threshold occur."""
    outs = Oracle(text=text).run(audio)[0]
    thr = ", ".join("%.15g" % np.quantile(outs[:, i], q) for i in range(outs.shape[1]))
    lines = [("thresholds = " + thr) if l.startswith("thresholds = ") else l for l in text.split("\n")]
    return "\n".join(lines)


if __name__ == "__main__":
    out = {}
    names = []
    for name, text, audio in cases():
        if name != "sample":
            text = retarget_thresholds(text, audio)
        o = Oracle(text=text)
        outs, da, df, band = o.run(audio, want_band=True)
        j0 = o.debounce(da, 0)
        j1 = o.debounce(da, o.debounce_frames(0.05))
        names.append(name)
        out[name + ".config"] = np.frombuffer(text.encode(), dtype=np.uint8)
        out[name + ".audio"] = audio
        out[name + ".outputs"] = outs
        out[name + ".det_any"] = da
        out[name + ".det_first"] = df
        out[name + ".band_head"] = band[:64]
        out[name + ".event_samples_d0"] = np.array([o.eval_sample(int(j)) for j in j0], dtype=np.int64)
        out[name + ".event_samples_d50ms"] = np.array([o.eval_sample(int(j)) for j in j1], dtype=np.int64)
        print(name, "evals", outs.shape, "detections", int(da.sum()), "events d0/d50ms", len(j0), len(j1))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_cases.npz"), **out)
