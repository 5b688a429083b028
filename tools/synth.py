"""Synthetic bird-song-like audio for tests and benchmarks (no datasets are reachable offline).

Audio = low-level Gaussian noise + a periodic "synthetic syllable": a bank of sinusoids at the band's bin centres whose
per-hop amplitudes follow a T x L template that drives the sample.txt network above threshold (the template comes from
tools/make_syllable.py; noise/tones/chirps alone never trigger that network).  Every syllable gets a seeded random gain
so both sides of the threshold occur.  The same recipe runs on numpy (host) and torch (device, for large corpora).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
TEMPLATE_PATH = os.path.join(os.path.dirname(_HERE), "tests", "golden", "syllable_template.npy")


def syllable_waveform(template, fft_len, hop, k0, amplitude=0.05):
    """-> float32 [T*hop] : sum_f template[t, f] * cos(2 pi (k0+f) n / N), row t held for one hop."""
    T, L = template.shape
    n = np.arange(T * hop)
    k = k0 + np.arange(L)
    amp = np.repeat(template.astype(np.float64), hop, axis=0)  # [T*hop, L]
    phase = 2.0 * np.pi * np.outer(n, k) / fft_len + 0.7 * np.arange(L)[None, :] ** 2
    return (amplitude * (amp * np.cos(phase)).sum(axis=1)).astype(np.float32)


def make_audio(n_channels, n_samples, seed=0, period_cols=40, hop=132, fft_len=256, k0=12, noise=1e-3, template=None):
    """numpy float32 [n_channels, n_samples]"""
    template = np.load(TEMPLATE_PATH) if template is None else template
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((n_channels, n_samples)) * noise).astype(np.float32)
    syl = syllable_waveform(template, fft_len, hop, k0)
    P = period_cols * hop
    n_per = n_samples // P
    if n_per > 0 and syl.size <= P:
        gains = rng.uniform(0.0, 1.6, size=(n_channels, n_per)).astype(np.float32)
        v = x[:, :n_per * P].reshape(n_channels, n_per, P)
        off = (7 * hop) // 3
        off = min(off, P - syl.size)
        v[:, :, off:off + syl.size] += gains[:, :, None] * syl[None, None, :]
    return x


def make_audio_torch(n_channels, n_samples, device, seed=0, period_cols=40, hop=132, fft_len=256, k0=12, noise=1e-3,
                     template=None):
    """torch float32 [n_channels, n_samples] generated on `device` (same recipe, torch's RNG)."""
    import torch

    template = np.load(TEMPLATE_PATH) if template is None else template
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.empty((n_channels, n_samples), dtype=torch.float32, device=device)
    x.normal_(0.0, noise, generator=g)
    syl = torch.from_numpy(syllable_waveform(template, fft_len, hop, k0)).to(device)
    P = period_cols * hop
    n_per = n_samples // P
    if n_per > 0 and syl.numel() <= P:
        gains = torch.empty((n_channels, n_per), dtype=torch.float32, device=device).uniform_(0.0, 1.6, generator=g)
        v = x[:, :n_per * P].view(n_channels, n_per, P)
        off = min((7 * hop) // 3, P - syl.numel())
        v[:, :, off:off + syl.numel()] += gains[:, :, None] * syl[None, None, :]
    return x
