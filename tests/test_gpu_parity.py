"""GPU parity tests (run with -m gpu on a B200). Everything goes through the C-ABI of libsyldet_cuda.so.

Tolerances (north_star: "spectra and network outputs within a stated FP32 tolerance, e.g. max rel err <= 1e-5; any
frame whose output lies within tolerance of the threshold reported separately"):
    network outputs   |gpu - oracle| <= TOL_OUT = 1e-5 absolute (linear scaling; db/log configs scale it, see below)
    detections        identical evaluation indices / sample numbers, except evaluations whose oracle output lies
                      within TOL_OUT of the threshold, which are counted and must be rare
"""
import importlib

import numpy as np
import pytest

from conftest import SAMPLE_TXT

pytestmark = pytest.mark.gpu
TOL_OUT = 1e-5
# db / log spectrogram scaling: the logarithm of a near-empty bin amplifies the float32 rounding of its magnitude (20 log10 of a bin
# at 1e-6 of the frame maximum moves by ~1e-3 per ulp-level change of the spectrum), identically in the oracle, its float64 twin and
# every kernel; network outputs of such configurations are compared at 2e-4 x output scale.
TOL_NONLINEAR = 2e-4
TOL_SPECTRA = 1e-5   # band magnitudes: |gpu - oracle| <= TOL_SPECTRA x the frame's largest band magnitude


@pytest.fixture(scope="module")
def cfg(sd):
    return sd.SyllableDetectorConfig(SAMPLE_TXT).validate()


@pytest.fixture(scope="module")
def orc(oracle_mod):
    return oracle_mod.Oracle(SAMPLE_TXT)


def _kernels(sd):
    return [(sd.KERNEL_TENSOR, "tensor"), (sd.KERNEL_FUSED, "fused"), (sd.KERNEL_GENERIC, "generic")]


def _check_events(orc, ref, outs, ev_samples, tol, debounce=0):
    """Detection parity without escape hatches. The GPU's decisions (its own outputs against the thresholds, compared in double like
    TrackDetector.swift:72) must equal the oracle's except at evaluations whose oracle output lies within `tol` of a threshold (those
    are returned, "reported separately"); and the GPU's event list must be exactly the greedy debounce (TrackDetector.swift:80,99) of
    its decisions - so with no flipped near-threshold evaluation the sample numbers are identical to the oracle's, debounce or not."""
    thr = orc.thresholds
    near = (np.abs(ref.astype(np.float64) - thr[None, :]) <= tol).any(axis=1)
    with np.errstate(invalid="ignore"):
        da_ref = (ref.astype(np.float64) >= thr[None, :]).any(axis=1)
        da_gpu = (outs.astype(np.float64) >= thr[None, :]).any(axis=1)
    flipped = da_ref != da_gpu
    assert not (flipped & ~near).any(), ("decision flips away from the threshold", np.nonzero(flipped & ~near)[0][:5])
    want = np.array([orc.eval_sample(int(j)) for j in orc.debounce(da_gpu, debounce)], dtype=np.int64)
    assert np.array_equal(ev_samples, want), "event list is not the debounce of the kernel's own decisions"
    if not flipped.any():
        ref_samples = np.array([orc.eval_sample(int(j)) for j in orc.debounce(da_ref, debounce)], dtype=np.int64)
        assert np.array_equal(ev_samples, ref_samples)
    return int(near.sum()), int(flipped.sum())


def _check_channel(orc, x, outs, ev_samples, tol, debounce=0):
    ref, da, _ = orc.run(x)
    assert outs.shape == ref.shape
    both_nan = np.isnan(outs) & np.isnan(ref)
    err = np.abs(np.where(both_nan, 0.0, outs - ref))
    assert not np.isnan(err).any() and err.max() <= tol, err.max()
    return _check_events(orc, ref, outs, ev_samples, tol, debounce)[0]


def test_device_present(sd):
    assert sd.device_count() >= 1


def test_sample_txt_parity_both_kernels(sd, cfg, orc, synth):
    x = synth.make_audio(3, 44100 * 6, seed=21)
    for kernel, name in _kernels(sd):
        det = sd.BatchDetector(cfg, kernel=kernel)
        assert det.active_kernel == kernel
        ev, outs = det.run(x, want_outputs=True)
        assert len(ev) > 20, name
        for ch in range(x.shape[0]):
            _check_channel(orc, x[ch], outs[ch], ev.sample[ev.channel == ch], TOL_OUT)
        assert np.allclose(ev.seconds, ev.sample / 44100.0)
        assert det.launch_count >= 1


def test_generic_kernel_spectra_are_bit_exact_for_linear_configs(sd, cfg, orc, synth):
    """The reference-order kernels reproduce the oracle's float32 arithmetic; with tanh the only libm call, outputs agree
    to a few ulp of the hidden activations."""
    x = synth.make_audio(1, 44100 * 2, seed=5)
    ev, outs = sd.BatchDetector(cfg, kernel=sd.KERNEL_GENERIC).run(x, want_outputs=True)
    ref = orc.run(x[0])[0]
    assert np.abs(outs[0] - ref).max() <= 1e-6


def test_golden_cases(sd, oracle_mod, golden):
    for name, g in golden.items():
        c = sd.SyllableDetectorConfig.from_text(g["config"]).validate()
        o = oracle_mod.Oracle(text=g["config"])
        scale = max(1.0, float(np.nanmax(np.abs(g["outputs"]))))
        tol = TOL_OUT * scale if c.spectrogram_scaling == "linear" else TOL_NONLINEAR * scale
        kernels = sd.BatchDetector.available_kernels(c)
        if name in ("sample", "log_std_128"):
            assert sd.KERNEL_FUSED in kernels, name  # these shapes must take a fast path
        if name == "sample":
            assert kernels[0] == sd.KERNEL_TENSOR
        for kernel in kernels:
            ev, outs = sd.BatchDetector(c, kernel=kernel).run(g["audio"], want_outputs=True)
            err = np.abs(outs[0] - g["outputs"])
            assert np.nanmax(err) <= tol, (name, kernel, float(np.nanmax(err)))
            near, flipped = _check_events(o, g["outputs"], outs[0], ev.sample, tol)
            D = c.debounce_frames(0.05)
            ev2 = sd.BatchDetector(c, kernel=kernel).run(g["audio"], debounce_frames=D)
            _check_events(o, g["outputs"], outs[0], ev2.sample, tol, debounce=D)
            if not flipped:   # then the committed event lists must be reproduced exactly
                assert np.array_equal(ev.sample, g["event_samples_d0"]), (name, kernel)
                assert np.array_equal(ev2.sample, g["event_samples_d50ms"]), (name, kernel)


@pytest.mark.parametrize("kw", [
    dict(fft_len=256, overlap=124, hidden=(4,), input_funcs=("l2normalize", "mapminmax")),
    dict(fft_len=256, win_len=200, overlap=60, hidden=(7,), input_funcs=("normalize", "mapminmax", "mapstd"), transfer="LogSig"),
    dict(fft_len=512, overlap=256, freq_range=(1000.0, 9000.0), time_range=5, hidden=(8, 5), outputs=3, input_funcs=("mapstd",), output_funcs=("mapminmax", "mapstd")),
    dict(fft_len=128, overlap=-7, freq_range=(300.0, 20000.0), time_range=12, hidden=(3,), input_funcs=("normalizestd", "mapminmax")),
    dict(fft_len=64, win_len=63, overlap=31, freq_range=(0.0, 22050.0), time_range=3, hidden=(), outputs=2, input_funcs=()),
    dict(fft_len=512, win_len=512, overlap=511, freq_range=(2000.0, 4000.0), time_range=20, hidden=(4,), input_funcs=("l2normalize",)),
    dict(fft_len=256, overlap=123, hidden=(4,), scaling="db", input_funcs=("mapminmax",)),
    dict(fft_len=1024, overlap=512, freq_range=(1000.0, 8000.0), time_range=4, hidden=(16,), input_funcs=("l2normalize", "mapminmax")),
    dict(fft_len=2048, win_len=1500, overlap=-100, freq_range=(500.0, 5000.0), time_range=2, hidden=(6,), scaling="log", input_funcs=("mapstd", "normalize")),
])
def test_generated_configs_fused_and_generic(sd, oracle_mod, cw, kw):
    text = cw.random_config(seed=7, threshold=0.3, **kw)
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    o = oracle_mod.Oracle(text=text)
    rng = np.random.default_rng(9)
    n = 40000
    t = np.arange(n)
    x = np.stack([(0.05 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * f0 * t / 44100 + 3 * np.sin(2 * np.pi * 2 * t / 44100))).astype(np.float32)
                  for f0 in (2500.0, 5200.0)])
    kernels = sd.BatchDetector.available_kernels(c)
    if kw["fft_len"] <= 512 and all(h <= 8 for h in kw.get("hidden", ())) and kw.get("input_funcs", ())[1:2] != ("normalize",):
        assert sd.KERNEL_FUSED in kernels
    for kernel in kernels:
        ev, outs = sd.BatchDetector(c, kernel=kernel).run(x, want_outputs=True)
        for ch in range(2):
            ref = o.run(x[ch])[0]
            scale = max(1.0, float(np.nanmax(np.abs(ref))))
            tol = (TOL_OUT if kw.get("scaling", "linear") == "linear" else TOL_NONLINEAR) * scale
            _check_channel(o, x[ch], outs[ch], ev.sample[ev.channel == ch], tol)


@pytest.mark.parametrize("kw", [
    # every instantiation / run-time branch of the tensor-core kernel besides the sample.txt shape (kFast):
    dict(overlap=124, hidden=(4,), input_funcs=("normalize", "mapminmax"), transfer="LogSig"),                  # min/max statistic partials
    dict(overlap=124, hidden=(4,), scaling="db", input_funcs=("mapminmax",)),                                    # kScaled, no statistic
    dict(overlap=128, hidden=(8,), time_range=6, outputs=2, input_funcs=("l2normalize",), output_funcs=("mapminmax", "mapstd")),  # HP = 8, several outputs
    dict(overlap=120, hidden=(4, 3), time_range=3, scaling="log", input_funcs=("l2normalize", "mapminmax")),     # three layers, hop 136, short window
    dict(overlap=124, hidden=(3,), time_range=12, freq_range=(500.0, 5900.0), input_funcs=("l2normalize", "mapminmax"), out_transfer="LogSig"),  # 32 bins, T = 12
    dict(overlap=124, hidden=(4,), input_funcs=("l2normalize", "mapminmax"), output_funcs=()),                  # sample shape without the reverse map
])
def test_tensor_kernel_shape_variants(sd, oracle_mod, cw, kw):
    text = cw.random_config(seed=11, threshold=0.3, fft_len=256, **kw)
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    assert sd.KERNEL_TENSOR in sd.BatchDetector.available_kernels(c), "shape must be eligible for the tensor-core kernel"
    o = oracle_mod.Oracle(text=text)
    rng = np.random.default_rng(13)
    n = 60000   # several tiles per unit, a ragged last tile
    t = np.arange(n)
    x = np.stack([(0.05 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * f0 * t / 44100 + 3 * np.sin(2 * np.pi * 2 * t / 44100))).astype(np.float32)
                  for f0 in (2500.0, 5200.0, 3900.0)])
    ev, outs = sd.BatchDetector(c, kernel=sd.KERNEL_TENSOR).run(x, want_outputs=True)
    for ch in range(x.shape[0]):
        ref = o.run(x[ch])[0]
        scale = max(1.0, float(np.nanmax(np.abs(ref))))
        tol = (TOL_OUT if kw.get("scaling", "linear") == "linear" else TOL_NONLINEAR) * scale
        _check_channel(o, x[ch], outs[ch], ev.sample[ev.channel == ch], tol)


def test_amplitude_invariance_default_kernel(sd, cfg, orc, synth):
    """The reference's float32 path is amplitude-invariant (l2normalize first: NeuralNet.swift:41-61). The default (AUTO = tensor)
    kernel runs two DFT correction products in fp16, guards the window in which that is exact enough on every evaluation and repeats
    the launch with its all-TF32 variant otherwise: outputs and events stay within the tolerance from rms 1e3 down to 2e-10, and the
    handle reports the switch. KERNEL_TENSOR_TF32 is the all-TF32 variant up front."""
    x0 = synth.make_audio(2, 44100 * 2, seed=5)
    for kernel in (sd.KERNEL_AUTO, sd.KERNEL_TENSOR, sd.KERNEL_TENSOR_TF32, sd.KERNEL_FUSED):
        for e in (20, 12, 0, -8, -16, -22):
            x = (x0 * np.float32(2.0 ** e)).astype(np.float32)
            det = sd.BatchDetector(cfg, kernel=kernel)
            ev, outs = det.run(x, want_outputs=True)
            for ch in range(x.shape[0]):
                _check_channel(orc, x[ch], outs[ch], ev.sample[ev.channel == ch], TOL_OUT)
            if kernel in (sd.KERNEL_AUTO, sd.KERNEL_TENSOR):
                assert det.active_kernel == sd.KERNEL_TENSOR
                assert det.range_fallbacks == (0 if e == 0 else 1 if e in (20, -16, -22) else det.range_fallbacks), (kernel, e)
            else:
                assert det.range_fallbacks == 0
    # the device-resident entry point settles the same way: collect() repeats the launch when the flag is up
    det = sd.BatchDetector(cfg)
    ev_q = det.run((x0 * np.float32(2.0 ** -20)).astype(np.float32))
    assert det.range_fallbacks == 1
    ev_1 = det.run(x0)                      # the handle stays on the all-TF32 variant: still correct
    s, _, _ = orc.events(x0[0], 0)
    assert np.array_equal(ev_1.sample[ev_1.channel == 0], s) and det.range_fallbacks == 1


def test_amplitude_int16_input_is_exact_for_the_fp16_pass(sd, cfg, orc, synth):
    """16-bit PCM k / 32768 gives exact fp16 operands whatever its level: even a recording of a few LSB stays on the fast variant."""
    rng = np.random.default_rng(12)
    for peak in (3, 40, 30000):
        s16 = np.clip(np.round(rng.standard_normal((2, 44100)) * peak / 3.0), -32768, 32767).astype(np.int16)
        xf = (s16.astype(np.float32) / 32768.0).astype(np.float32)
        det = sd.BatchDetector(cfg)
        ev, outs = det.run(s16, want_outputs=True)
        assert det.active_kernel == sd.KERNEL_TENSOR and det.range_fallbacks == 0
        for ch in range(2):
            _check_channel(orc, xf[ch], outs[ch], ev.sample[ev.channel == ch], TOL_OUT)


def test_amplitude_range_non_normalised_configs(sd, oracle_mod, cw):
    """Configurations without a per-window normaliser (mapminmax only; db scaling) are not scale-invariant: the tensor kernel keeps all
    three DFT products in TF32 for them, so no amplitude triggers a fallback and the error stays at float32 level. "float32 level" is
    amplitude-dependent here (magnitudes of ~30 against offsets of ~5e-4 make the hidden pre-activations ill-conditioned): the bound is
    the larger of the usual tolerance and 4x the distance between the float32 oracle and its float64 twin on the same input."""
    from oracle import twin64
    for kw in (dict(hidden=(4,), input_funcs=("mapminmax",)), dict(hidden=(4,), scaling="db", input_funcs=("mapminmax",))):
        text = cw.random_config(seed=17, threshold=0.3, fft_len=256, overlap=124, **kw)
        c = sd.SyllableDetectorConfig.from_text(text).validate()
        o = oracle_mod.Oracle(text=text)
        tw = twin64.Twin64(text=text)
        rng = np.random.default_rng(4)
        n = 50000
        t = np.arange(n)
        x0 = (0.05 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * 3100.0 * t / 44100)).astype(np.float32)
        for e in (-16, 0, 8, 18):
            x = (x0 * np.float32(2.0 ** e)).astype(np.float32)
            det = sd.BatchDetector(c, kernel=sd.KERNEL_TENSOR)
            ev, outs = det.run(x, want_outputs=True)
            ref = o.run(x)[0]
            r64 = tw.run(x)
            r64 = r64[0] if isinstance(r64, tuple) else r64
            scale = max(1.0, float(np.nanmax(np.abs(ref))))
            tol = (TOL_OUT if kw.get("scaling", "linear") == "linear" else TOL_NONLINEAR) * scale
            tol = max(tol, 4.0 * float(np.nanmax(np.abs(ref - r64))))
            _check_channel(o, x, outs[0], ev.sample, tol)
            assert det.range_fallbacks == 0


def test_amplitude_invariance_minmax_window_config(sd, oracle_mod, cw):
    """A window normalised by its own minimum / maximum ("normalize") is scale-invariant too; the tensor kernel runs such shapes on
    the fp16-split band DFT as well, guarded by the window's range max - min instead of its norm: quiet input (range < 2^-8) and
    fp16 overflow make the handle repeat the launch with the all-TF32 variant, everything in between stays on the fast one."""
    text = cw.random_config(seed=23, threshold=0.3, fft_len=256, overlap=128, freq_range=(1500.0, 6500.0), time_range=6, hidden=(8,),
                            input_funcs=("normalize", "mapminmax"), transfer="LogSig")
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    o = oracle_mod.Oracle(text=text)
    rng = np.random.default_rng(8)
    n = 60000
    t = np.arange(n)
    x0 = (0.05 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 2900.0 * t / 44100) * (np.sin(2 * np.pi * 3.0 * t / 44100) > 0)).astype(np.float32)
    for e, fallbacks in ((0, 0), (6, 0), (-6, 0), (-20, 1), (20, 1)):
        x = (x0 * np.float32(2.0 ** e)).astype(np.float32)
        det = sd.BatchDetector(c, kernel=sd.KERNEL_TENSOR)
        ev, outs = det.run(x, want_outputs=True)
        assert det.active_kernel == sd.KERNEL_TENSOR
        scale = max(1.0, float(np.nanmax(np.abs(o.run(x)[0]))))
        _check_channel(o, x, outs[0], ev.sample, TOL_OUT * scale)
        assert det.range_fallbacks == fallbacks, (e, det.range_fallbacks)


@pytest.mark.parametrize("kernel_name", ["KERNEL_TENSOR", "KERNEL_TENSOR_TF32", "KERNEL_FUSED", "KERNEL_GENERIC"])
def test_spectra_match_oracle(sd, cfg, orc, synth, kernel_name):
    """extractPower()[f0 ..< f1] (CSTFT.swift:280-337) from every kernel against the oracle's float32 radix-2 FFT: north_star's
    "spectra ... within a stated FP32 tolerance", here |delta| <= 1e-5 x the frame's largest band magnitude."""
    x = synth.make_audio(3, 132 * 700 + 1444 + 57, seed=27)
    x[2] *= 300.0
    det = sd.BatchDetector(cfg, kernel=getattr(sd, kernel_name))
    band = det.spectra(x)
    E = cfg.num_evals(x.shape[1])
    assert band.shape == (3, E + cfg.time_range - 1, 29)
    for ch in range(3):
        ref = orc.stft_band(x[ch])[:band.shape[1]]
        frame_max = ref.max(axis=1, keepdims=True)
        assert (frame_max > 0).all()
        assert (np.abs(band[ch] - ref) / frame_max).max() <= TOL_SPECTRA, (kernel_name, ch, float((np.abs(band[ch] - ref) / frame_max).max()))
    # the tap does not disturb the detection path
    ev, outs = det.run(x, want_outputs=True)
    _check_channel(orc, x[0], outs[0], ev.sample[ev.channel == 0], TOL_OUT)


@pytest.mark.parametrize("kw", [
    dict(fft_len=256, overlap=124, hidden=(4,), scaling="db", input_funcs=("mapminmax",)),                       # tensor (kScaled) + fused + generic
    dict(fft_len=128, overlap=-7, freq_range=(300.0, 20000.0), time_range=12, hidden=(3,), input_funcs=("l2normalize",)),   # gap: fused + generic
    dict(fft_len=512, win_len=400, overlap=256, freq_range=(1000.0, 9000.0), time_range=5, hidden=(8,), scaling="log", input_funcs=("mapstd",)),   # zero padding
])
def test_spectra_generated_configs(sd, oracle_mod, cw, kw):
    """Spectra are reported before the db / log scaling (they are extractPower's values), for every kernel the shape qualifies for."""
    text = cw.random_config(seed=5, threshold=0.3, **kw)
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    o = oracle_mod.Oracle(text=text)
    rng = np.random.default_rng(6)
    n = 30000
    t = np.arange(n)
    x = (0.05 * rng.standard_normal(n) + 0.4 * np.sin(2 * np.pi * 2900.0 * t / 44100 + 2 * np.sin(2 * np.pi * 3 * t / 44100))).astype(np.float32)
    lin = oracle_mod.Oracle(text=text.replace("scaling = %s" % kw.get("scaling", "linear"), "scaling = linear"))
    ref_all = lin.stft_band(x)
    for kernel in sd.BatchDetector.available_kernels(c):
        band = sd.BatchDetector(c, kernel=kernel).spectra(x)[0]
        ref = ref_all[:band.shape[0]]
        assert band.shape[0] == o.num_evals(n) + c.time_range - 1
        frame_max = ref.max(axis=1, keepdims=True)
        assert (np.abs(band - ref) / frame_max).max() <= TOL_SPECTRA, (kernel, float((np.abs(band - ref) / frame_max).max()))


def test_high_overlap_wide_hidden_config(sd, oracle_mod, cw):
    """BASELINE config 4 shape: FFT 1024, hop 4, band 1-8 kHz (L = 162), T = 8 (1296 inputs), 256 tansig units, 2 outputs.
    Too wide for the fused kernels: AUTO takes the wide-hidden tensor path (layer 0 as a 3xTF32 tcgen05 contraction over the sliding
    feature window, kernels_wide.cu); the reference-order kernels remain selectable and agree with it."""
    text = cw.random_config(seed=21, fft_len=1024, overlap=1020, freq_range=(1000.0, 8000.0), time_range=8, hidden=(256,),
                            outputs=2, threshold=0.2)
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    assert (c.hop, c.net_inputs) == (4, 1296)
    o = oracle_mod.Oracle(text=text)
    rng = np.random.default_rng(3)
    n = 1024 + 4 * 7 + 4 * 700
    t = np.arange(n)
    x = np.stack([(0.05 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * f0 * t / 44100)).astype(np.float32) for f0 in (3000.0, 6100.0)])
    det = sd.BatchDetector(c)
    assert det.active_kernel == sd.KERNEL_WIDE
    assert sd.BatchDetector.available_kernels(c) == [sd.KERNEL_WIDE, sd.KERNEL_GENERIC]
    for kernel in (sd.KERNEL_WIDE, sd.KERNEL_GENERIC):
        ev, outs = sd.BatchDetector(c, kernel=kernel).run(x, want_outputs=True)
        assert outs.shape[1] == 701
        for ch in range(2):
            ref = o.run(x[ch])[0]
            scale = max(1.0, float(np.nanmax(np.abs(ref))))
            _check_channel(o, x[ch], outs[ch], ev.sample[ev.channel == ch], TOL_OUT * scale)


@pytest.mark.parametrize("kw", [
    dict(fft_len=1024, overlap=1020, freq_range=(1000.0, 8000.0), time_range=8, hidden=(600,), outputs=1, input_funcs=("normalize", "mapminmax"), transfer="LogSig"),   # 3 accumulator passes, min/max statistic
    dict(fft_len=512, win_len=400, overlap=396, freq_range=(500.0, 9000.0), time_range=5, hidden=(64,), outputs=4, scaling="db", input_funcs=("mapminmax",), output_funcs=("mapminmax", "mapstd")),   # zero padding, no statistic, 4 outputs
    dict(fft_len=256, overlap=252, freq_range=(2000.0, 7000.0), time_range=17, hidden=(1024,), outputs=2, input_funcs=("l2normalize",), out_transfer="TanSig"),   # T at the limit, 4 passes, 29 bins (2 chunks, padded)
])
def test_wide_kernel_shape_variants(sd, oracle_mod, cw, kw):
    """Every branch of the wide-hidden tensor path: several accumulator passes (hidden > 256), each window statistic, spectrogram
    scaling, plane padding, ragged tiles (evaluation counts that are no multiple of 256) and several channels."""
    text = cw.random_config(seed=31, threshold=0.3, **kw)
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    assert c.hop == 4 and sd.KERNEL_WIDE in sd.BatchDetector.available_kernels(c)
    o = oracle_mod.Oracle(text=text)
    rng = np.random.default_rng(17)
    n = kw.get("win_len", kw["fft_len"]) + 4 * (kw["time_range"] - 1) + 4 * 1100 + 3
    t = np.arange(n)
    x = np.stack([(0.05 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * f0 * t / 44100 + 2 * np.sin(2 * np.pi * 5 * t / 44100))).astype(np.float32)
                  for f0 in (3000.0, 6100.0, 4400.0)])
    ev, outs = sd.BatchDetector(c, kernel=sd.KERNEL_WIDE).run(x, want_outputs=True)
    assert outs.shape[1] == 1101
    for ch in range(x.shape[0]):
        ref = o.run(x[ch])[0]
        scale = max(1.0, float(np.nanmax(np.abs(ref))))
        tol = (TOL_OUT if kw.get("scaling", "linear") == "linear" else TOL_NONLINEAR) * scale
        _check_channel(o, x[ch], outs[ch], ev.sample[ev.channel == ch], tol)
    band = sd.BatchDetector(c, kernel=sd.KERNEL_WIDE).spectra(x[:1])
    assert band.shape == (1, 1101 + kw["time_range"] - 1, o.L)


@pytest.mark.parametrize("kernel_name", ["KERNEL_FUSED", "KERNEL_TENSOR", "KERNEL_GENERIC"])
def test_chunk_and_batch_invariance(sd, cfg, synth, kernel_name):
    """Any split of the work gives identical results: channels together or alone, long or short recordings.
    The tensor-core kernel hands the last time_range + 1 evaluations of a buffer (whose hop-rows are incomplete) to the SIMT
    kernel, so there the match is bit-exact up to that tail and within TOL_OUT inside it."""
    kernel = getattr(sd, kernel_name)
    tail = cfg.time_range + 1 if kernel == sd.KERNEL_TENSOR else 0

    def same(a, b):
        assert a.shape == b.shape
        n = a.shape[-2] - tail
        assert np.array_equal(a[..., :n, :], b[..., :n, :])
        assert np.abs(a - b).max() <= TOL_OUT

    x = synth.make_audio(5, 132 * 3000 + 77, seed=33)
    det = sd.BatchDetector(cfg, kernel=kernel)
    ev_all, out_all = det.run(x, want_outputs=True)
    for ch in (0, 4):
        ev1, out1 = det.run(x[ch], want_outputs=True)
        assert np.array_equal(out1[0], out_all[ch]) and np.array_equal(ev1.sample, ev_all.sample[ev_all.channel == ch])
    # prefix property: outputs of a prefix equal the prefix of the outputs
    n2 = 132 * 1000 + 1444
    ev2, out2 = det.run(x[:, :n2], want_outputs=True)
    same(out2, out_all[:, :out2.shape[1]])
    # time-shift by whole hops shifts evaluations (pure function of the sample span)
    ev3, out3 = det.run(x[:, 132 * 17:], want_outputs=True)
    same(out3, out_all[:, 17:])


def test_edge_sizes(sd, cfg, orc):
    det = sd.BatchDetector(cfg)
    rng = np.random.default_rng(0)
    for n in (0, 1, 255, 1443, 1444, 1445, 1444 + 131, 1444 + 132, 1444 + 132 * 255, 1444 + 132 * 256 + 5):
        x = (rng.standard_normal((2, max(n, 1))) * 0.01).astype(np.float32)[:, :n] if n else np.zeros((2, 0), dtype=np.float32)
        if n == 0:
            continue  # a zero-length buffer is rejected upstream too (guard 0 < numSamples, TrackDetector.swift:53)
        ev, outs = det.run(x, want_outputs=True)
        assert outs.shape[1] == orc.num_evals(n)
        for ch in range(2):
            if outs.shape[1]:
                assert np.abs(outs[ch] - orc.run(x[ch])[0]).max() <= TOL_OUT
    # silence: NaN outputs, no detection
    ev, outs = det.run(np.zeros((1, 5000), dtype=np.float32), want_outputs=True)
    assert np.isnan(outs).all() and len(ev) == 0


def test_debounce_and_detect_rules(sd, oracle_mod, cw, synth, cfg, orc):
    x = synth.make_audio(2, 44100 * 5, seed=8)
    det = sd.BatchDetector(cfg)
    for seconds in (0.0, 0.01, 0.05, 0.5):
        D = cfg.debounce_frames(seconds)
        ev = det.run(x, debounce_frames=D)
        for ch in range(2):
            s, _, _ = orc.events(x[ch], D)
            assert np.array_equal(ev.sample[ev.channel == ch], s)
    # any-output (CLI) vs first-output (live) rule on a 2-output network
    text = cw.random_config(seed=12, hidden=(4,), outputs=2, threshold=[10.0, -10.0])
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    o = oracle_mod.Oracle(text=text)
    d2 = sd.BatchDetector(c)
    assert len(d2.run(x[0], detect_rule=sd.DETECT_ANY_OUTPUT)) == o.num_evals(x.shape[1])
    assert len(d2.run(x[0], detect_rule=sd.DETECT_FIRST_OUTPUT)) == 0


def test_event_buffer_overflow_replay(sd, cw):
    """Every evaluation detects: more events than the initial device buffer holds -> grow and replay."""
    text = cw.random_config(seed=2, hidden=(4,), threshold=-1e9)
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    x = (np.random.default_rng(0).standard_normal((12, 132 * 100000 + 1444)) * 0.01).astype(np.float32)
    ev = sd.BatchDetector(c).run(x)
    assert len(ev) == 12 * c.num_evals(x.shape[1]) > (1 << 20)
    assert np.array_equal(ev.sample[:3], [1444, 1576, 1708]) and np.all(np.diff(ev.channel) >= 0)
    # the same through the 16-bit path that feeds the tensor kernel directly (sample network, threshold far below every output): the
    # repeat needs the float32 copy of the recording that path had skipped
    text = open(SAMPLE_TXT).read().replace("threshold = 0.442442442442442", "threshold = -1000000000.0")
    assert "threshold = -1000000000.0" in text
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    s16 = np.clip(np.round(x * 32768.0 * 8.0), -32768, 32767).astype(np.int16)
    det = sd.BatchDetector(c)
    ev = det.run(s16)
    assert det.active_kernel == sd.KERNEL_TENSOR
    assert len(ev) == 12 * c.num_evals(x.shape[1]) > (1 << 20)
    assert np.array_equal(ev.sample[:3], [1444, 1576, 1708]) and np.all(np.diff(ev.channel) >= 0)


def test_run_into_event_table_matches_run(sd, cfg, synth):
    """BatchDetector.run_into: the library writes a recording's detections straight into a sharding.EventTable as compact gather rows -
    the same rows pack_events_compact builds from run()'s columns, order tracked across recordings."""
    sh = importlib.import_module("syllable-detector-swift_b200.sharding")
    det = sd.BatchDetector(cfg)
    table = sh.EventTable(cfg.net_outputs, capacity=8)
    want = []
    for rec, (nch, n, seed) in enumerate([(3, 44100 * 4, 41), (1, 44100 * 2, 42), (2, 5000, 43), (4, 44100 * 3, 44)]):
        x = synth.make_audio(nch, n, seed=seed)
        ev = det.run(x)
        assert det.run_into(table, rec, x) == len(ev)
        want.append(sh.pack_events_compact(rec, ev.channel, ev.sample, ev.outputs))
    want = np.concatenate(want)
    assert table.n == want.shape[0] > 20 and table.in_order and np.array_equal(table.rows, want)
    x = synth.make_audio(2, 44100 * 2, seed=45)
    det.run_into(table, 1, x)                      # an earlier recording after a later one: the table notices
    assert not table.in_order and sh.rows_in_order(sh.gather_events(table, None))


def test_pcm16_and_interleaved_ingest(sd, cfg, synth):
    x = synth.make_audio(3, 44100 * 2, seed=4) * 8.0
    s16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    xf = (s16.astype(np.float32) / 32768.0).astype(np.float32)
    det = sd.BatchDetector(cfg)
    ev_ref, out_ref = det.run(xf, want_outputs=True)
    ev_a, out_a = det.run(s16, want_outputs=True)                                             # planar int16
    ev_b, out_b = det.run(np.ascontiguousarray(xf.T), want_outputs=True, layout=sd.LAYOUT_INTERLEAVED)   # interleaved float
    ev_c, out_c = det.run(np.ascontiguousarray(s16.T), want_outputs=True, layout=sd.LAYOUT_INTERLEAVED)  # interleaved int16
    for o, e in ((out_a, ev_a), (out_b, ev_b), (out_c, ev_c)):
        assert np.array_equal(o, out_ref) and np.array_equal(e.sample, ev_ref.sample) and np.array_equal(e.channel, ev_ref.channel)


@pytest.mark.parametrize("kernel_name", ["KERNEL_TENSOR", "KERNEL_FUSED", "KERNEL_GENERIC"])
def test_run_host_time_slices_are_invisible(sd, cfg, orc, synth, kernel_name):
    """run() pipelines the recording in time slices (copy of slice k+1 over detection of slice k): events, outputs and the
    debounce must not depend on the slicing, for every input format; one channel is checked against the oracle."""
    x = synth.make_audio(3, 44100 * 6 + 77, seed=21) * 4.0
    s16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    xq = (s16.astype(np.float32) / 32768.0).astype(np.float32)
    det = sd.BatchDetector(cfg, kernel=getattr(sd, kernel_name))
    det.set_slice_evals(1 << 40)                       # one slice
    ev1, out1 = det.run(xq, want_outputs=True, debounce_frames=700)
    assert len(ev1) > 10
    _check_channel(orc, xq[1], out1[1], ev1.sample[ev1.channel == 1], 1e-5, debounce=700)
    for slice_evals in (600, 5000):                    # 16 slices (the cap), then 1 slice per ~5000 evaluations
        det.set_slice_evals(slice_evals)
        for pcm, layout in ((xq, sd.LAYOUT_PLANAR), (s16, sd.LAYOUT_PLANAR), (np.ascontiguousarray(xq.T), sd.LAYOUT_INTERLEAVED),
                            (np.ascontiguousarray(s16.T), sd.LAYOUT_INTERLEAVED)):
            ev, out = det.run(pcm, want_outputs=True, debounce_frames=700, layout=layout)
            assert np.array_equal(out, out1)
            assert np.array_equal(ev.sample, ev1.sample) and np.array_equal(ev.channel, ev1.channel)
            assert np.array_equal(ev.outputs, ev1.outputs)
    # a strided planar source (channel_stride > n_samples is what a caller with padded rows passes) and no events at all
    det.set_slice_evals(600)
    ev0 = det.run(np.zeros((2, 44100), np.float32))
    assert len(ev0) == 0


def test_syllable_detector_api_matches_oracle(sd, cfg, orc, synth):
    """class SyllableDetector: appendAudioData / processNewValue / lastOutputs / lastDetected with ragged buffers."""
    x = synth.make_audio(1, 44100 * 2, seed=14)[0]
    ref, da, df = orc.run(x)
    d = sd.SyllableDetector(cfg)
    rng = np.random.default_rng(1)
    pos, got, flags = 0, [], []
    while pos < x.size:
        n = int(rng.choice([1, 32, 100, 131, 132, 133, 500, 4096]))
        d.append_audio_data(x[pos:pos + n])
        pos += n
        while d.process_new_value():
            got.append(d.last_outputs.copy())
            flags.append(d.last_detected)
    got = np.array(got)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= TOL_OUT
    assert np.array_equal(np.array(flags), df) and df.sum() > 0
    # seenSyllable(): OR over everything pending
    d2 = sd.SyllableDetector(cfg)
    d2.append_audio_data(x[:44100])
    assert d2.seen_syllable() == bool(df[:orc.num_evals(44100)].any())
    assert d2.process_new_value() is False
    # ring overflow is an error, not an abort (CSTFT.swift:199)
    d3 = sd.SyllableDetector(cfg)
    d3.append_audio_data(np.zeros(102400, dtype=np.float32))
    with pytest.raises(sd.SyldetError) as e:
        d3.append_audio_data(np.zeros(1, dtype=np.float32))
    assert e.value.kind == "bufferOverflow"


def test_track_detector_rows(sd, cfg, orc, synth):
    x = synth.make_audio(1, 44100 * 3, seed=15)[0]
    D = cfg.debounce_frames(0.02)
    bufs = [x[i:i + 8192] for i in range(0, x.size, 8192)]
    td = sd.TrackDetector(bufs, cfg, channel=3)
    td.debounce_time = 0.02
    assert td.debounce_frames == D
    while td.process():
        pass
    s, sec, outs = orc.events(x, D)
    assert [r[1] for r in td.rows] == list(s) and all(r[0] == 3 for r in td.rows)
    assert np.allclose([r[2] for r in td.rows], sec) and np.abs(np.array([r[3] for r in td.rows]) - outs).max() <= TOL_OUT


def test_stream_group_matches_batch(sd, cfg, orc, synth):
    """Processor.swift shape: 16 channels, 32-frame buffers; per-tick `seen` flags and outputs equal the offline run."""
    nch, nbuf, ticks = 16, 32, 1500
    x = synth.make_audio(nch, nbuf * ticks, seed=16)
    g = sd.StreamGroup(cfg, nch, max_buffer=nbuf)
    total_new = np.zeros(nch, dtype=np.int64)
    seen_eval = [[] for _ in range(nch)]
    last = None
    for t in range(ticks):
        seen, n_new = g.submit(x[:, t * nbuf:(t + 1) * nbuf])
        for ch in range(nch):
            if n_new[ch]:
                seen_eval[ch].append((int(total_new[ch]), int(n_new[ch]), bool(seen[ch])))
        total_new += n_new
        last = g.last_outputs.copy()
    E = orc.num_evals(nbuf * ticks)
    assert np.all(total_new == E)
    for ch in (0, 7, 15):
        ref, _, df = orc.run(x[ch])
        for start, cnt, flag in seen_eval[ch]:
            assert flag == bool(df[start:start + cnt].any())
        assert np.abs(last[ch] - ref[-1]).max() <= TOL_OUT


@pytest.mark.parametrize("kw", [
    dict(fft_len=128, overlap=-7, freq_range=(300.0, 20000.0), time_range=12, hidden=(3,), input_funcs=("normalizestd", "mapminmax")),
    dict(fft_len=512, overlap=256, freq_range=(1000.0, 9000.0), time_range=5, hidden=(8, 5), outputs=3, input_funcs=("mapstd",), output_funcs=("mapminmax", "mapstd")),
    dict(fft_len=2048, win_len=1500, overlap=-100, freq_range=(500.0, 5000.0), time_range=2, hidden=(6,), scaling="log", input_funcs=("mapstd", "normalize")),
    # shapes the latency-shaped tick (stream_tick_fast_kernel) takes: min/max window statistic + zero padding + gap; FFT 128 with dB
    # scaling and 8 hidden units; FFT 64 with the l2 statistic
    dict(fft_len=256, win_len=200, overlap=-5, freq_range=(1000.0, 8000.0), time_range=6, hidden=(4,), input_funcs=("normalize", "mapminmax")),
    dict(fft_len=128, overlap=64, freq_range=(1500.0, 12000.0), time_range=7, hidden=(8,), outputs=2, scaling="db", input_funcs=("mapminmax",)),
    dict(fft_len=64, overlap=32, freq_range=(2000.0, 15000.0), time_range=8, hidden=(4,), input_funcs=("l2normalize",)),
])
def test_stream_group_ragged_ticks_generated_configs(sd, oracle_mod, cw, kw):
    """The live tick kernel (device sample ring + band-feature ring, one launch per tick) on configurations with a gap, several
    outputs and long windows, fed ragged buffer lengths (1 sample .. several hops: single-launch and per-phase launches)."""
    text = cw.random_config(seed=11, threshold=0.3, **kw)
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    o = oracle_mod.Oracle(text=text)
    rng = np.random.default_rng(5)
    nch, n = 3, 30000
    x = (0.1 * rng.standard_normal((nch, n))).astype(np.float32)
    refs = [o.run(x[ch])[0] for ch in range(nch)]
    scale = max(1.0, max(float(np.nanmax(np.abs(r))) for r in refs))
    tol = (TOL_OUT if kw.get("scaling", "linear") == "linear" else TOL_NONLINEAR) * scale
    g = sd.StreamGroup(c, nch, max_buffer=3000)
    pos, done, launches = 0, 0, 0
    while pos < n:
        m = min(int(rng.choice([1, 7, 32, 64, 257, 1000, 3000])), n - pos)
        seen, n_new = g.submit(x[:, pos:pos + m])
        pos += m
        assert np.all(n_new == n_new[0])
        if n_new[0]:
            done += int(n_new[0])
            for ch in range(nch):
                assert np.abs(g.last_outputs[ch] - refs[ch][done - 1]).max() <= tol
                flag = bool((refs[ch][done - int(n_new[0]):done, 0].astype(np.float64) >= c.thresholds[0]).any())
                near = bool((np.abs(refs[ch][done - int(n_new[0]):done, 0].astype(np.float64) - c.thresholds[0]) <= tol).any())
                assert near or bool(seen[ch]) == flag
    assert done == o.num_evals(n) > 0
    assert 0 < g.launch_count
    if kw["fft_len"] <= 256 and "normalizestd" not in kw.get("input_funcs", ()):   # shapes of the latency-shaped tick: it did take them
        assert g.fast_tick_count > 0


@pytest.mark.parametrize("rate_in", [0.0, 48000.0])
def test_stream_group_resident_kernel(sd, cfg, orc, synth, monkeypatch, rate_in):
    """SYLDET_STREAM_RESIDENT=1: the tick blocks stay on their SMs and poll a message in pinned host memory. Same results, tick by tick
    and bit for bit, as the launched tick - across an idle period longer than the kernel's idle limit (it leaves and is started again),
    a level-meter read (which stops it) and, second case, with the resampler inside the tick."""
    import time
    nch, nbuf, ticks = 12, 32, 900
    x = synth.make_audio(nch, nbuf * ticks, seed=23)
    kw = dict(max_buffer=nbuf)
    if rate_in:
        kw["input_rate"] = rate_in
    ref_group = sd.StreamGroup(cfg, nch, **kw)
    monkeypatch.setenv("SYLDET_STREAM_RESIDENT", "1")
    monkeypatch.setenv("SYLDET_STREAM_RESIDENT_IDLE_MS", "5")
    g = sd.StreamGroup(cfg, nch, **kw)
    monkeypatch.delenv("SYLDET_STREAM_RESIDENT")
    total = 0
    for t in range(ticks):
        buf = x[:, t * nbuf:(t + 1) * nbuf]
        seen_a, new_a = ref_group.submit(buf)
        seen_b, new_b = g.submit(buf)
        assert np.array_equal(new_a, new_b) and np.array_equal(seen_a, seen_b)
        if new_a[0]:
            assert np.array_equal(ref_group.last_outputs, g.last_outputs)
            total += int(new_a[0])
        if t == 300:
            time.sleep(0.05)          # ten idle limits: the kernel has left; the next tick starts it again
        if t == 601:                  # one buffer is waiting in the staging area: the meter read sends it through the resident kernel first
            la, lb = ref_group.read_levels(), g.read_levels()
            assert np.allclose(la[0], lb[0], rtol=1e-6, equal_nan=True) and np.array_equal(la[1], lb[1], equal_nan=True)
    assert total > 0 and g.resident_tick_count > 0 and g.resident_tick_count >= g.launch_count
    assert ref_group.resident_tick_count == 0
    if not rate_in:
        ref, _, _ = orc.run(x[3])
        assert np.abs(g.last_outputs[3] - ref[total - 1]).max() <= TOL_OUT


def test_stream_group_resident_kernel_ragged_buffers(sd, cfg, synth, monkeypatch):
    """Ragged buffers on a resident group: ticks of up to four columns stay on the resident kernel (several staged buffers per message),
    longer buffers stop it, go through the per-phase launches and the next short tick starts it again - every output bit for bit what a
    launched group gives."""
    nch, n = 5, 120000
    x = synth.make_audio(nch, n, seed=29)
    ref_group = sd.StreamGroup(cfg, nch, max_buffer=4000)
    monkeypatch.setenv("SYLDET_STREAM_RESIDENT", "1")
    g = sd.StreamGroup(cfg, nch, max_buffer=4000)
    monkeypatch.delenv("SYLDET_STREAM_RESIDENT")
    rng = np.random.default_rng(31)
    pos, done = 0, 0
    while pos < n:
        m = min(int(rng.choice([1, 7, 32, 32, 32, 64, 132, 257, 500, 700, 4000])), n - pos)
        buf = x[:, pos:pos + m]
        seen_a, new_a = ref_group.submit(buf)
        seen_b, new_b = g.submit(buf)
        pos += m
        assert np.array_equal(new_a, new_b) and np.array_equal(seen_a, seen_b)
        if new_a[0]:
            assert np.array_equal(ref_group.last_outputs, g.last_outputs)
            done += int(new_a[0])
    assert done == cfg.num_evals(n) and g.resident_tick_count > 100 and g.launch_count > 10


def test_stream_level_meters_and_pulses(sd, cfg, orc, synth):
    """Live view extras: input RMS / output maximum meters (Processor.swift:110-113, 138, 158-184) and the 1 ms TTL pulse train
    (Processor.swift:192, 212-221; AudioInterface.swift:13-40, 442-445) over 8 channels of 32-frame buffers."""
    nch, nbuf, ticks = 8, 32, 1400
    x = synth.make_audio(nch, nbuf * ticks, seed=31)
    g = sd.StreamGroup(cfg, nch, max_buffer=nbuf)
    g.set_pulse(0.001, 44100.0)
    a, b = g.read_levels()
    assert np.isnan(a).all() and np.isnan(b).all()          # nil before anything arrived
    refs = [orc.run(x[ch]) for ch in range(nch)]
    done, t_prev, e_prev = 0, 0, 0
    pulses = []
    for t in range(ticks):
        seen, n_new = g.submit(x[:, t * nbuf:(t + 1) * nbuf])
        done += int(n_new[0])
        p = g.render_pulses(nbuf)
        pulses.append(p)
        if (t + 1) % 350 == 0 or t == 6:                     # t == 6: buffers still waiting in the staging area are metered too
            a, b = g.read_levels()
            seg = x[:, t_prev * nbuf:(t + 1) * nbuf].reshape(nch, -1, nbuf)
            ms = (seg.astype(np.float32) ** 2).sum(axis=2, dtype=np.float32).astype(np.float64) / nbuf
            assert np.allclose(a, np.sqrt(ms.max(axis=1)), rtol=1e-6, atol=0)
            for ch in range(nch):
                if done > e_prev:
                    assert abs(b[ch] - float(np.nanmax(refs[ch][0][e_prev:done, 0]))) <= TOL_OUT
                else:
                    assert np.isnan(b[ch])
            t_prev, e_prev = t + 1, done
    # pulse train: armed to Int(0.001 * 44100) = 44 frames by every tick with a detection; rendered 32 frames per callback
    pulses = np.concatenate(pulses, axis=1)
    assert set(np.unique(pulses)) <= {0.0, 1.0}
    for ch in (0, 5):
        df = refs[ch][2]
        high, exp, e = 0, [], 0
        for t in range(ticks):
            new = orc.num_evals((t + 1) * nbuf) - e
            if new and df[e:e + new].any():
                high = 44
            e += new
            exp.append(np.arange(nbuf) < high)
            high -= min(high, nbuf)
        assert np.array_equal(pulses[ch] > 0.5, np.concatenate(exp)) and pulses[ch].sum() > 0


@pytest.mark.parametrize("rate_in,nbuf", [(48000.0, 32), (96000.0, 100), (22050.0, 32), (44100.5, 32)])
def test_stream_group_resamples_inside_the_tick(sd, cfg, orc, oracle_mod, synth, rate_in, nbuf):
    """Processor.swift:116-121: every 32-frame buffer goes through ResamplerLinear before appendAudioData. The group fed at the
    device rate must behave exactly like the oracle's resampler (bit-for-bit samples, hence the usual output tolerance) followed by
    the detector - including the buffer-size dependence and the one-sample-per-buffer carry of the upstream resampler."""
    nch, ticks = 5, 1200
    x = synth.make_audio(nch, nbuf * ticks, seed=51)           # interpreted as audio at rate_in
    g = sd.StreamGroup(cfg, nch, max_buffer=nbuf, input_rate=rate_in)
    assert g.resampling == (abs(rate_in - 44100.0) > 1.0)
    rs = [oracle_mod.Resampler(rate_in, 44100.0) for _ in range(nch)]
    y = [[] for _ in range(nch)]
    got_new = np.zeros(nch, dtype=np.int64)
    checked = 0
    for t in range(ticks):
        buf = x[:, t * nbuf:(t + 1) * nbuf]
        for ch in range(nch):
            y[ch].append(rs[ch].process(buf[ch]) if g.resampling else buf[ch])
        seen, n_new = g.submit(buf)
        got_new += n_new
        if n_new[0] and t >= 97 * (checked + 1):            # spot checks along the way: the newest evaluation of every channel
            for ch in range(nch):
                ref = orc.run(np.concatenate(y[ch]))[0]
                assert ref.shape[0] == got_new[ch]
                assert np.abs(g.last_outputs[ch] - ref[-1]).max() <= TOL_OUT
            checked += 1
    assert checked > 2
    for ch in range(nch):
        yc = np.concatenate(y[ch])
        ref, _, df = orc.run(yc)
        assert got_new[ch] == ref.shape[0] > 100
    # the meters see the device-rate samples (Processor.swift:110-113 runs before the resampler)
    a, _ = g.read_levels()
    seg = x.reshape(nch, ticks, nbuf)
    ms = (seg.astype(np.float32) ** 2).sum(axis=2, dtype=np.float32).astype(np.float64) / nbuf
    assert np.allclose(a, np.sqrt(ms.max(axis=1)), rtol=1e-6, atol=0)


def test_simulator_trace_matches_oracle(sd, cfg, orc, oracle_mod, synth):
    """Simulator output track (ViewControllerSimulator.swift:251-254, 308-344): float trace within TOL_OUT / thr0 of the oracle's,
    16-bit trace within two quantisation steps; exact structure (leading zeros, hop-long plateaus)."""
    n = 44100 * 3 + 77
    x = synth.make_audio(2, n, seed=23)
    det = sd.BatchDetector(cfg)
    tr = det.simulate(x)
    tr16 = det.simulate(x, s16=True)
    first, hop, thr0 = cfg.first_output_sample, cfg.hop, cfg.thresholds[0]
    assert tr.shape == (2, n) and tr16.dtype == np.int16
    for ch in range(2):
        ref_out = orc.run(x[ch])[0][:, 0]
        ref = oracle_mod.simulator_trace(ref_out, thr0, first, hop, n)
        ref16 = oracle_mod.simulator_trace(ref_out, thr0, first, hop, n, s16=True)
        assert np.all(tr[ch, :first] == 0.0) and np.abs(tr[ch] - ref).max() <= TOL_OUT / thr0 * 1.01
        E = orc.num_evals(n)
        body = tr[ch, first:first + (E - 1) * hop].reshape(E - 1, hop)   # the last evaluation's plateau is cut short by the end
        assert np.all(body == body[:, :1]) and np.all(tr[ch, first + (E - 1) * hop:] == tr[ch, first + (E - 1) * hop])
        assert tr[ch].min() >= 0.0 and tr[ch].max() == 1.0   # the synthetic syllables saturate the trace
        assert np.abs(tr16[ch].astype(np.int32) - ref16.astype(np.int32)).max() <= 2 and tr16[ch].max() == 32767
    # silence: l2normalize gives NaN outputs; the float trace keeps them, the 16-bit trace stores 0
    z = np.zeros((1, 4000), dtype=np.float32)
    assert np.isnan(det.simulate(z)[0, first:first + hop]).all() and np.all(det.simulate(z, s16=True) == 0)


def test_resampler_matches_oracle_bit_for_bit(sd, oracle_mod):
    rng = np.random.default_rng(2)
    x = rng.standard_normal(32 * 400).astype(np.float32)
    for rin, rout, nbuf in ((48000, 44100, 32), (44100, 48000, 32), (96000, 44100, 512), (22050, 44100, 100), (44100, 44100, 64)):
        a, b = sd.ResamplerLinear(rin, rout), oracle_mod.Resampler(rin, rout)
        for i in range(0, x.size - nbuf + 1, nbuf):
            ya, yb = a.resample_vector(x[i:i + nbuf]), b.process(x[i:i + nbuf])
            assert ya.size == yb.size and np.array_equal(ya, yb), (rin, rout, i)


def test_batched_linear_resampler_matches_oracle_bit_for_bit(sd, oracle_mod):
    """Whole-channel ResamplerLinear.resampleVector (Resampler.swift:35-70) on the device, several channels per launch, against the
    oracle's single call from a fresh state - including the float32 index ramp's loss of precision on long buffers."""
    rng = np.random.default_rng(8)
    for rin, rout, n in ((48000, 44100, 300001), (96000, 44100, 123457), (22050, 44100, 70000), (44100, 48000, 50003), (44100, 44100, 4097), (48000, 44100, 3)):
        x = rng.standard_normal((3, n)).astype(np.float32)
        y = sd.resample(x, rin, rout, mode=sd.RESAMPLE_LINEAR)
        for ch in range(3):
            ref = oracle_mod.Resampler(rin, rout).process(x[ch])
            assert y[ch].size == ref.size and np.array_equal(y[ch], ref), (rin, rout, ch)


def test_polyphase_resampler_matches_scipy(sd):
    """The quality converter (north_star (1)): same design and alignment as scipy.signal.resample_poly; float32 taps and accumulation
    against scipy's float64: |delta| <= 1e-5 x the signal's peak. Also: a tone stays a tone (nothing else above -70 dB)."""
    import scipy.signal as ss
    rng = np.random.default_rng(9)
    for rin, rout, n in ((48000, 44100, 100003), (96000, 44100, 88211), (22050, 44100, 30000), (44100, 48000, 44100), (32000, 44100, 16001), (48000, 44100, 5)):
        x = rng.standard_normal((2, n)).astype(np.float32)
        y = sd.resample(x, rin, rout)
        g = np.gcd(rin, rout)
        ref = ss.resample_poly(x.astype(np.float64), rout // g, rin // g, axis=1)
        assert y.shape == ref.shape, (rin, rout, y.shape, ref.shape)
        assert np.abs(y - ref).max() <= 1e-5 * np.abs(x).max(), (rin, rout, float(np.abs(y - ref).max()))
    t = np.arange(96000) / 48000.0
    tone = np.sin(2 * np.pi * 3000.0 * t).astype(np.float32)
    y = sd.resample(tone, 48000, 44100)[2000:-2000]
    spec = np.abs(np.fft.rfft(y * np.hanning(y.size)))
    k = int(round(3000.0 * y.size / 44100.0))
    spec[k - 16:k + 17] = 0.0
    assert 20 * np.log10(spec.max() / (y.size / 4.0)) < -70.0   # the kaiser(5.0) design's stop band, as scipy's
    with pytest.raises(sd.SyldetError):
        sd.resample(tone, 44100.5, 44100)


def test_cli_other_rates_and_sample_formats(sd, cfg, synth, tmp_path):
    """`syldet` on a 48 kHz 24-bit file: converted to the network's rate on the device (upstream: AVFoundation delivers
    config.samplingRate, SyllableDetector.swift:19-23); on a 44.1 kHz 24-bit file: uploaded as packed integers (SYLDET_PCM_S24)."""
    import os
    import subprocess
    from conftest import ROOT, SAMPLE_TXT as NET
    exe = os.path.join(ROOT, "syllable-detector-swift_b200", "syldet")

    def write24(path, x, rate):
        q = np.clip(np.round(x * 8388608.0), -8388608, 8388607).astype(np.int32)          # [ch, n]
        b = np.ascontiguousarray(q.T).astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3].tobytes()
        nch = x.shape[0]
        hdr = b"RIFF" + np.uint32(36 + len(b)).tobytes() + b"WAVEfmt " + np.uint32(16).tobytes() + np.uint16(1).tobytes() + np.uint16(nch).tobytes() + \
            np.uint32(rate).tobytes() + np.uint32(rate * nch * 3).tobytes() + np.uint16(nch * 3).tobytes() + np.uint16(24).tobytes() + b"data" + np.uint32(len(b)).tobytes()
        with open(path, "wb") as f:
            f.write(hdr + b)
        return (q.astype(np.float32) / np.float32(8388608.0)).astype(np.float32)

    def rows_of(path):
        out = subprocess.run([exe, "-n", NET, "-a", path], capture_output=True, text=True, check=True).stdout
        return [(int(r[0]), int(r[1]), float(r[3])) for r in (l.split(",") for l in out.strip().splitlines())]

    x = synth.make_audio(2, 44100 * 3, seed=29) * 6.0
    xq = write24(str(tmp_path / "a24.wav"), x, 44100)
    det = sd.BatchDetector(cfg)
    ev = det.run(xq)
    got = rows_of(str(tmp_path / "a24.wav"))
    assert len(got) == len(ev) > 5
    assert [g[:2] for g in got] == list(zip(ev.channel.tolist(), ev.sample.tolist()))
    assert np.abs(np.array([g[2] for g in got]) - ev.outputs[:, 0]).max() <= 1e-6
    # the same material played at 48 kHz (resampled up by scipy in float64, stored as 24-bit): the CLI converts it back on the device
    import scipy.signal as ss
    x48 = ss.resample_poly(x.astype(np.float64), 160, 147, axis=1).astype(np.float32)
    x48q = write24(str(tmp_path / "a48.wav"), x48, 48000)
    back = sd.resample(x48q, 48000, 44100)
    ev48 = det.run(back)
    got48 = rows_of(str(tmp_path / "a48.wav"))
    assert len(got48) == len(ev48) > 5 and [g[:2] for g in got48] == list(zip(ev48.channel.tolist(), ev48.sample.tolist()))
    # and the detections agree with the original's, up to evaluations that sit near the threshold (two conversions lie in between)
    assert abs(len(ev48) - len(ev)) <= max(3, len(ev) // 20)


def test_large_run_properties(sd, cfg, orc, synth):
    """BASELINE config-2 shape at reduced length per channel but full channel count: size-independent checks."""
    nch, n = 8, 44100 * 120
    x = synth.make_audio(nch, n, seed=40)
    det = sd.BatchDetector(cfg)
    ev, outs = det.run(x, want_outputs=True)
    E = cfg.num_evals(n)
    assert outs.shape == (nch, E, 1) and not np.isnan(outs).any()
    # events == thresholded dense outputs, sorted by (channel, sample), samples on the hop lattice
    flags = outs[:, :, 0].astype(np.float64) >= cfg.thresholds[0]
    assert len(ev) == int(flags.sum()) > 1000
    assert np.all((ev.sample - 1444) % 132 == 0)
    key = ev.channel.astype(np.int64) * (1 << 40) + ev.sample
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(outs[ev.channel, (ev.sample - 1444) // 132, 0], ev.outputs[:, 0])
    # sampled slices against the oracle
    rng = np.random.default_rng(3)
    for _ in range(6):
        ch, j = int(rng.integers(nch)), int(rng.integers(E - 300))
        seg = x[ch, j * 132: j * 132 + 1444 + 132 * 299]
        assert np.abs(outs[ch, j:j + 300] - orc.run(seg)[0]).max() <= TOL_OUT
    # all kernels agree everywhere
    for k in (sd.KERNEL_GENERIC, sd.KERNEL_FUSED, sd.KERNEL_TENSOR):
        ev_g, outs_g = sd.BatchDetector(cfg, kernel=k).run(x[:2], want_outputs=True)
        assert np.abs(outs_g - outs[:2]).max() <= TOL_OUT, k


def test_cli_csv_rows(sd, cfg, orc, synth, tmp_path):
    """`syldet -n net -a wav -d s` prints channel,sample,seconds,out0 like SyllableDetectorCLI (main.swift:31-39)."""
    import os
    import subprocess
    import wave
    from conftest import ROOT, SAMPLE_TXT as NET
    x = synth.make_audio(2, 44100 * 3, seed=19) * 8.0
    s16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(np.ascontiguousarray(s16.T).tobytes())
    exe = os.path.join(ROOT, "syllable-detector-swift_b200", "syldet")
    out = subprocess.run([exe, "-n", NET, "-a", path, "-d", "0.02"], capture_output=True, text=True, check=True).stdout
    rows = [l.split(",") for l in out.strip().splitlines()]
    xf = s16.astype(np.float32) / 32768.0
    want = []
    for ch in range(2):
        s, sec, outs = orc.events(xf[ch], cfg.debounce_frames(0.02))
        want += [(ch, int(a), float(b), float(c[0])) for a, b, c in zip(s, sec, outs)]
    assert len(rows) == len(want) > 5
    for r, (ch, s, sec, o) in zip(rows, want):
        assert int(r[0]) == ch and int(r[1]) == s and abs(float(r[2]) - sec) < 1e-12 and abs(float(r[3]) - o) <= TOL_OUT
    assert subprocess.run([exe], capture_output=True).returncode == 64  # EX_USAGE (main.swift:40)
