"""Builds tests/golden/syllable_template.npy: a non-negative T x L magnitude template that drives the sample.txt
network above threshold, found by seeded coordinate hill-climbing on the float64 twin (SURVEY.md section 8(d) recipe).
Noise, tones and chirps never trigger this network, so benchmark/test audio mixes in this "synthetic syllable".
Run from the repo root:  python tools/make_syllable.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.twin64 import Twin64  # noqa: E402


def climb(t, seed=1, iters=6000):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.05, 1.0, t.I)
    best = t.net(x[None, :])[0, 0]
    for _ in range(iters):
        y = x.copy()
        idx = rng.integers(0, t.I, size=rng.integers(1, 6))
        y[idx] = np.clip(y[idx] * np.exp(rng.normal(0, 0.5, idx.size)), 1e-3, 10.0)
        v = t.net(y[None, :])[0, 0]
        if v > best:
            x, best = y, v
    return x.reshape(t.T, t.L), best


if __name__ == "__main__":
    t = Twin64(os.path.join(ROOT, "tests", "golden", "sample.txt"))
    tpl, best = climb(t)
    tpl = tpl / tpl.max()
    print("template output", best, "threshold", t.thr)
    np.save(os.path.join(ROOT, "tests", "golden", "syllable_template.npy"), tpl.astype(np.float32))
