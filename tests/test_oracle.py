"""CPU tests of the oracle itself: analytic known-answer tests, float64 twin, golden regression, reference quirks.
(The reference has no tests or golden vectors for this path - SURVEY.md section 4 - so these anchor the restatement.)"""
import numpy as np
import pytest

from conftest import SAMPLE_TXT


@pytest.fixture(scope="module")
def orc(oracle_mod):
    return oracle_mod.Oracle(SAMPLE_TXT)


def test_sample_txt_geometry(orc):
    # SURVEY Appendix A / section 8(c) KATs
    assert (orc.fourierLength, orc.windowLength, orc.windowOverlap) == (256, 256, 124)
    assert (orc.stride, orc.gap, orc.k0, orc.k1, orc.L, orc.I, orc.O) == (132, 0, 12, 41, 29, 290, 1)
    assert orc.eval_sample(0) == 1444 and orc.eval_sample(7) == 1444 + 7 * 132
    assert orc.num_columns(2646000) == 20044 and orc.num_evals(2646000) == 20035
    assert orc.array("layer0.weights").size == 1160 and orc.array("layer0.biases").size == 4
    assert orc.array("processInputs1.xOffsets").size == 290 and orc.array("processInputs1.gains").size == 290
    assert abs(orc.thresholds[0] - 0.442442442442442) < 1e-15


def test_window_kats(oracle_mod, orc):
    w = orc.array("window")
    assert abs(w.sum() - 0.54 * 256) < 1e-4  # sum of N-denominator Hamming = 0.54 N
    assert abs(w[0] - 0.08) < 1e-7 and abs(w[128] - 1.0) < 1e-7
    from oracle.pyoracle import make_window
    assert np.allclose(make_window(2, 64), 0.5 * (1 - np.cos(2 * np.pi * np.arange(64) / 64)), atol=1e-7)
    assert np.allclose(make_window(3, 64), 0.42 - 0.5 * np.cos(2 * np.pi * np.arange(64) / 64) + 0.08 * np.cos(4 * np.pi * np.arange(64) / 64), atol=1e-7)
    assert np.all(make_window(0, 16) == 1.0)


def test_spectrum_kats(orc):
    n = np.arange(256)
    w = orc.array("window")
    # impulse at n0 -> |X[k]| = w[n0] for every k
    imp = np.zeros(256, dtype=np.float32)
    imp[37] = 1.0
    assert np.allclose(orc.frame_spectrum(imp), w[37], rtol=2e-6)
    # DC c -> |X[0]| = c * 0.54 N
    assert abs(orc.frame_spectrum(np.full(256, 0.25, dtype=np.float32))[0] - 0.25 * 0.54 * 256) < 1e-4
    # sinusoid at exact bin k0: A*0.54*N/2 at k0, A*0.23*N/2 at k0 +- 1
    s = orc.frame_spectrum((0.5 * np.cos(2 * np.pi * 20 * n / 256)).astype(np.float32))
    assert abs(s[20] - 0.5 * 0.54 * 128) < 1e-4 and abs(s[19] - 0.5 * 0.23 * 128) < 1e-4 and abs(s[21] - 0.5 * 0.23 * 128) < 1e-4
    # linearity in amplitude is exact for powers of two
    x = np.random.default_rng(0).standard_normal(256).astype(np.float32)
    assert np.array_equal(orc.frame_spectrum(x * 4.0), orc.frame_spectrum(x) * 4.0)


def test_oracle_matches_float64_twin(oracle_mod, synth):
    from oracle.twin64 import Twin64
    o, t = oracle_mod.Oracle(SAMPLE_TXT), Twin64(SAMPLE_TXT)
    x = synth.make_audio(1, 44100 * 3, seed=2)[0]
    o32, da, _, band = o.run(x, want_band=True)
    b64, o64 = t.band(x), t.run(x)
    assert band.shape == b64.shape and o32.shape == o64.shape
    assert (np.abs(band - b64).max(axis=1) / b64.max(axis=1)).max() < 1e-6   # relative to the per-frame max
    assert np.abs(o32 - o64).max() < 1e-5
    near = np.abs(o64[:, 0] - t.thr[0]) < 1e-5
    assert np.array_equal(da[~near], (o64[:, 0] >= t.thr[0])[~near]) and da.sum() > 0
    assert np.array_equal(t.eval_sample(np.arange(5)), [o.eval_sample(j) for j in range(5)])


def test_golden_regression(oracle_mod, golden):
    for name, g in golden.items():
        o = oracle_mod.Oracle(text=g["config"])
        outs, da, df, band = o.run(g["audio"], want_band=True)
        assert np.array_equal(outs, g["outputs"], equal_nan=True), name
        assert np.array_equal(da, g["det_any"].astype(bool)) and np.array_equal(df, g["det_first"].astype(bool)), name
        assert np.array_equal(band[:64], g["band_head"]), name
        s0, _, _ = o.events(g["audio"], 0)
        s1, sec, eo = o.events(g["audio"], o.debounce_frames(0.05))
        assert np.array_equal(s0, g["event_samples_d0"]) and np.array_equal(s1, g["event_samples_d50ms"]), name
        assert np.allclose(sec, s1 / o.samplingRate)


def test_generated_configs_match_twin(oracle_mod, cw, golden):
    from oracle.twin64 import Twin64
    for name, g in golden.items():
        o, t = oracle_mod.Oracle(text=g["config"]), Twin64(text=g["config"])
        assert (o.k0, o.k1, o.stride, o.gap, o.I, o.O) == (t.k0, t.k1, t.stride, t.gap, t.I, t.O)
        a = o.run(g["audio"])[0]
        b = t.run(g["audio"])
        scale = max(1.0, np.nanmax(np.abs(b)))
        assert np.nanmax(np.abs(a - b)) < 2e-4 * scale, name  # db/log scaling amplifies float32 error of small bins


def test_counts_and_edges(oracle_mod, orc):
    assert orc.num_columns(255) == 0 and orc.num_columns(256) == 1 and orc.num_columns(256 + 131) == 1 and orc.num_columns(256 + 132) == 2
    assert orc.num_evals(1443) == 0 and orc.num_evals(1444) == 1 and orc.num_evals(1444 + 131) == 1 and orc.num_evals(1444 + 132) == 2
    outs, da, df = orc.run(np.zeros(1000, dtype=np.float32))
    assert outs.shape == (0, 1)
    # silence: l2normalize divides 0 by 0 -> NaN -> never detected
    outs, da, df = orc.run(np.zeros(3000, dtype=np.float32))
    assert outs.shape[0] == orc.num_evals(3000) and np.all(np.isnan(outs)) and not da.any() and not df.any()
    # trailing partial hop is dropped (no flush at EOF)
    x = np.random.default_rng(1).standard_normal(5000).astype(np.float32)
    a = orc.run(x)[0]
    b = orc.run(x[:1444 + 132 * (a.shape[0] - 1)])[0]
    assert np.array_equal(a, b)


def test_chunk_invariance_of_oracle(orc, synth):
    x = synth.make_audio(1, 30000, seed=3)[0]
    full = orc.run(x)[0]
    # evaluation j depends only on samples [j*s, j*s + W + (T-1)*s)
    j = 57
    seg = x[j * 132: j * 132 + 1444]
    assert np.array_equal(orc.run(seg)[0][0], full[j])


def test_debounce_rule(orc):
    det = np.zeros(100, dtype=bool)
    det[[3, 4, 5, 20, 21, 60]] = True
    assert list(orc.debounce(det, 0)) == [3, 4, 5, 20, 21, 60]
    # emit iff until < S_j ; until = S_j + D.  D = 2 hops -> j=4 (S+132), j=5 (S+264 == until) suppressed
    assert list(orc.debounce(det, 264)) == [3, 20, 60]
    assert list(orc.debounce(det, 263)) == [3, 5, 20, 60]
    assert orc.debounce_frames(0.05) == 2205 and orc.debounce_frames(0.0499999) == 2204


def test_net_apply_direct(oracle_mod, orc):
    from oracle.twin64 import Twin64
    t = Twin64(SAMPLE_TXT)
    v = np.random.default_rng(4).uniform(0.01, 1.0, 290).astype(np.float32)
    assert abs(orc.net_apply(v)[0] - t.net(v[None, :].astype(np.float64))[0, 0]) < 1e-5
    # mapminmax reverse of sample.txt is (y + 1) / 2
    assert orc.array("processOutputs0.gains")[0] == 2.0 and orc.scalar("processOutputs0.y") == -1.0


def test_resampler_oracle(oracle_mod):
    R = oracle_mod.Resampler
    x = np.sin(np.arange(4096) * 0.01).astype(np.float32)
    # identity rate: first buffer passes through
    assert np.array_equal(R(44100, 44100).process(x[:32]), x[:32])
    # downsample 48k -> 44.1k in one buffer: y[k] = lerp(x, k*step)
    r = R(48000, 44100)
    y = r.process(x[:512])
    step = np.float32(48000 / 44100)
    idx = (np.arange(y.size, dtype=np.float32) * step).astype(np.float32)
    b = idx.astype(np.int64)
    ref = x[b] + (idx - b.astype(np.float32)) * (x[b + 1] - x[b])
    assert y.size == int(np.float32(512) / step) and np.array_equal(y, ref.astype(np.float32))
    # stateful and segmentation dependent: ~ (n-1)/n of the ideal count with 32-frame buffers
    r = R(48000, 44100)
    n_out = sum(r.process(x[i:i + 32]).size for i in range(0, 4096, 32))
    ideal = 4096 * 44100 / 48000
    assert 0.955 < n_out / ideal < 0.98
    st = r.state()
    assert abs(st[0] - step) == 0 and -1.0 <= st[2] <= 1.0


def test_simulator_trace_known_answer():
    """ViewControllerSimulator.swift:251-254, 326-344: silent until the first evaluation is due, then clamp(out0 / thr0, 0, 1)
    held for one hop; the 16-bit form saturates at 32767 and stores NaN as 0."""
    import oracle
    tr = oracle.simulator_trace([0.25, 1.0, -0.1, np.nan], 0.5, first=5, hop=3, n_samples=16)
    exp = np.array([0, 0, 0, 0, 0, .5, .5, .5, 1, 1, 1, 0, 0, 0, np.nan, np.nan], dtype=np.float32)
    assert np.array_equal(tr, exp, equal_nan=True)
    tr16 = oracle.simulator_trace([0.25, 1.0, -0.1, np.nan], 0.5, first=5, hop=3, n_samples=16, s16=True)
    assert tr16.tolist() == [0] * 5 + [16384] * 3 + [32767] * 3 + [0] * 5
    # more evaluations than the recording has room for: the trace stops at n_samples
    assert oracle.simulator_trace([1.0] * 10, 1.0, first=2, hop=4, n_samples=9).tolist() == [0, 0, 1, 1, 1, 1, 1, 1, 1]
