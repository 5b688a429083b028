"""convert_to_text.m counterpart and the text-format linter. CPU only: parsing goes through the C-ABI's
config functions, which need no GPU."""
import importlib
import os
import warnings

import numpy as np
import pytest

from conftest import SAMPLE_TXT


@pytest.fixture(scope="module")
def ct(sd):
    return importlib.import_module("syllable-detector-swift_b200.convert_to_text")


def _mat_dict_from_config(c, as_cells=True):
    """The struct convert_to_text.m expects (convert_to_text.m:30-35, 61-116), filled from a parsed configuration."""
    n = c.layer_count
    layers = [c.layer(i) for i in range(n)]
    tf_names = {"TanSig": "tansig", "LogSig": "logsig", "PureLin": "purelin", "SatLin": "satlin"}

    def put(procs):
        fcns, settings = [], []
        for fn, xo, g, y in procs:
            if fn in ("mapminmax", "mapstd"):
                fcns.append(fn)
                settings.append({"xoffset": xo.astype(np.float64), "gain": g.astype(np.float64), "ymin" if fn == "mapminmax" else "ymean": float(y)})
        return {"processFcns": np.array(fcns, dtype=object), "processSettings": np.array(settings, dtype=object)}

    IW = np.empty((n, 1), dtype=object)
    LW = np.empty((n, n), dtype=object)
    B = np.empty((n, 1), dtype=object)
    for i in range(n):
        w, b, tf = layers[i]
        IW[i, 0] = w.astype(np.float64) if i == 0 else np.zeros((0, 0))
        for j in range(n):
            LW[i, j] = w.astype(np.float64) if (i > 0 and j == i - 1) else np.zeros((0, 0))
        B[i, 0] = b.astype(np.float64).reshape(-1, 1)
    net = {"input": put(c.input_processing), "output": put(c.output_processing),
           "layers": np.array([{"netInputFcn": "netsum", "transferFcn": tf_names[layers[i][2]]} for i in range(n)], dtype=object),
           "IW": IW, "LW": LW, "b": B}
    return {"samplerate": c.sampling_rate, "fft_size": c.fourier_length, "win_size": c.window_length,
            "fft_time_shift": c.fourier_length - c.window_overlap, "freq_range": np.array(c.freq_range), "time_window_steps": c.time_range,
            "trigger_thresholds": np.array(c.thresholds), "scaling": c.spectrogram_scaling, "net": net}


def _same(a, b):
    assert (a.sampling_rate, a.fourier_length, a.window_length, a.window_overlap, a.time_range) == \
           (b.sampling_rate, b.fourier_length, b.window_length, b.window_overlap, b.time_range)
    assert a.freq_range == b.freq_range and a.spectrogram_scaling == b.spectrogram_scaling
    assert np.array_equal(a.thresholds, b.thresholds) and a.layer_count == b.layer_count
    for i in range(a.layer_count):
        wa, ba, ta = a.layer(i)
        wb, bb, tb = b.layer(i)
        assert np.array_equal(wa, wb) and np.array_equal(ba, bb) and ta == tb
    for pa, pb in ((a.input_processing, b.input_processing), (a.output_processing, b.output_processing)):
        assert [p[0] for p in pa] == [p[0] for p in pb]
        for (fa, xa, ga, ya), (fb, xb, gb, yb) in zip(pa, pb):
            if xa is not None:
                assert np.array_equal(xa, xb) and np.array_equal(ga, gb) and ya == yb


def test_sample_txt_round_trips_through_mat(sd, ct, tmp_path):
    """sample.txt -> the exporter's .mat struct (saved and re-loaded by scipy) -> convert_to_text -> identical configuration.
    l2normalize is not a Matlab processFcn: it travels as 'prepend_input_processing', exactly as upstream does it."""
    from scipy.io import savemat
    c = sd.SyllableDetectorConfig(SAMPLE_TXT).validate()
    d = _mat_dict_from_config(c)
    path = os.path.join(tmp_path, "net.mat")
    savemat(path, d)
    out = os.path.join(tmp_path, "net.txt")
    text = ct.convert_to_text(out, path, prepend_input_processing="l2normalize")
    assert open(out).read() == text and text.startswith("# AUTOMATICALLY GENERATED SYLLABLE DETECTOR CONFIGURATION\nsamplingRate = 44100.0\n")
    c2 = sd.SyllableDetectorConfig.from_text(text).validate()
    _same(c, c2)
    assert "thresholds = " in text and [ln for ln in ct.lint_config_text(text)] == []
    # the dict form gives the same bytes as the .mat form
    assert ct.convert_to_text(None, d, prepend_input_processing=("l2normalize",)) == text


def test_generated_network_round_trip_and_checks(sd, cw, ct):
    text = cw.random_config(seed=3, fft_len=512, overlap=256, freq_range=(1000.0, 9000.0), time_range=5, hidden=(8, 5), outputs=3,
                            input_funcs=("mapstd",), output_funcs=("mapminmax", "mapstd"), threshold=[0.1, 0.2, 0.3])
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    d = _mat_dict_from_config(c)
    _same(c, sd.SyllableDetectorConfig.from_text(ct.convert_to_text(None, d)).validate())
    bad = dict(d, fft_size=300)
    with pytest.raises(ct.ConvertError, match="power of two"):
        ct.convert_to_text(None, bad)
    with pytest.raises(ct.ConvertError, match="window size"):
        ct.convert_to_text(None, dict(d, win_size=1024))
    net = dict(d["net"])
    net["layers"] = np.array([{"netInputFcn": "netsum", "transferFcn": "radbas"}] * 3, dtype=object)
    with pytest.raises(ct.ConvertError, match="Invalid transfer function"):
        ct.convert_to_text(None, dict(d, net=net))
    with warnings.catch_warnings(record=True) as w:     # convert_to_text.m:50-53: FFT sizes below 256 are raised to 256
        warnings.simplefilter("always")
        small = dict(d, fft_size=128, win_size=128, fft_time_shift=64)
        t = ct.convert_to_text(None, small)
    assert "fourierLength = 256\n" in t and "windowOverlap = 192\n" in t and any("256" in str(x.message) for x in w)


def test_linter_reports_the_parsers_silent_cases(ct):
    text = open(SAMPLE_TXT).read()
    assert ct.lint_config_text(text) == ["line 8: legacy key 'threshold' (still accepted, SyllableDetectorConfig.swift:223-229)"]
    noisy = text + "note = a = b\nlayer7.inputs = 3\nfooBar = 1\nscaling = db\n# kept = as key\nprocessInputs5.function = mapminmax\n"
    msgs = "\n".join(ct.lint_config_text(noisy))
    for needle in ("does not have the form key = value", "layers = 2 but keys of later layers", "'fooBar' is not read", "repeats line",
                   "starts with '#'", "processInputsCount = 2 but entries up to index 5"):
        assert needle in msgs, (needle, msgs)
