"""Writer for the SyllableDetectorConfig text format.

Emits the same keys, order and number formatting (`%.15g`, weights row-major [outputs x inputs]) as the
reference's Matlab exporter (convert_to_text.m:61-74 header, :118-182 processing functions, :184-212
layers), so that files written here load in the reference app and vice versa.  Used to build synthetic
networks (seeded random weights) for the configurations BASELINE.json names.
"""
import math

import numpy as np

TRANSFER = ("TanSig", "LogSig", "PureLin", "SatLin")
INPUT_FUNCS = ("mapminmax", "mapstd", "l2normalize", "normalize", "normalizestd")


def _g(v):
    return "%.15g" % float(v)


def _arr(a):
    return ", ".join(_g(v) for v in np.asarray(a, dtype=np.float64).ravel())


def freq_index_range(fft_len, f_lo, f_hi, fs):
    """CircularShortTimeFourierTransform.frequencyIndexRange (CSTFT.swift:166-191)."""
    half = fft_len // 2
    frm = fft_len / fs
    k0 = int(math.ceil(frm * f_lo))
    k1 = min(int(math.floor(frm * f_hi)) + 1, half)
    return k0, k1


def write_config(sampling_rate, fft_len, win_len, overlap, freq_range, time_range, thresholds, scaling, layers,
                 process_inputs=(), process_outputs=(), legacy_threshold_key=False):
    """layers: [(W [out, in], b [out], transfer name)], process_*: [(function, xOffsets, gains, y)] -> str"""
    out = ["# AUTOMATICALLY GENERATED SYLLABLE DETECTOR CONFIGURATION",
           "samplingRate = %.1f" % sampling_rate,
           "fourierLength = %d" % fft_len,
           "windowLength = %d" % win_len,
           "windowOverlap = %d" % overlap,
           "freqRange = %.1f, %.1f" % (freq_range[0], freq_range[1]),
           "timeRange = %d" % time_range,
           "%s = %s" % ("threshold" if legacy_threshold_key else "thresholds", _arr(thresholds)),
           "scaling = %s" % scaling]

    def procs(prefix, items):
        out.append("%sCount = %d" % (prefix, len(items)))
        for k, (fn, xo, g, y) in enumerate(items):
            out.append("%s%d.function = %s" % (prefix, k, fn))
            if fn in ("mapminmax", "mapstd"):
                out.append("%s%d.xOffsets = %s" % (prefix, k, _arr(xo)))
                out.append("%s%d.gains = %s" % (prefix, k, _arr(g)))
                out.append("%s%d.%s = %s" % (prefix, k, "yMin" if fn == "mapminmax" else "yMean", _g(y)))

    procs("processInputs", list(process_inputs))
    procs("processOutputs", list(process_outputs))
    out.append("layers = %d" % len(layers))
    for i, (w, b, tf) in enumerate(layers):
        w = np.asarray(w)
        assert tf in TRANSFER and w.ndim == 2 and w.shape[0] == np.asarray(b).size
        out += ["layer%d.inputs = %d" % (i, w.shape[1]), "layer%d.outputs = %d" % (i, w.shape[0]),
                "layer%d.weights = %s" % (i, _arr(w)), "layer%d.biases = %s" % (i, _arr(b)),
                "layer%d.transferFunction = %s" % (i, tf)]
    return "\n".join(out) + "\n"


def random_config(seed=0, fs=44100.0, fft_len=256, win_len=None, overlap=124, freq_range=(2000.0, 7000.0),
                  time_range=10, hidden=(4,), outputs=1, scaling="linear", input_funcs=("l2normalize", "mapminmax"),
                  output_funcs=("mapminmax",), transfer="TanSig", out_transfer="PureLin", threshold=0.5):
    """Seeded random network of the given shape, written in the reference text format."""
    rng = np.random.default_rng(seed)
    win_len = fft_len if win_len is None else win_len
    k0, k1 = freq_index_range(fft_len, freq_range[0], freq_range[1], fs)
    n_in = (k1 - k0) * time_range
    dims = [n_in] + list(hidden) + [outputs]
    layers = []
    for i in range(len(dims) - 1):
        w = rng.standard_normal((dims[i + 1], dims[i])) / math.sqrt(dims[i]) * (3.0 if i == 0 else 1.5)
        b = rng.standard_normal(dims[i + 1]) * 0.5
        layers.append((w, b, transfer if i < len(dims) - 2 else out_transfer))

    def proc(fn, n, is_input):
        if fn == "mapminmax":
            if is_input:
                return (fn, rng.uniform(0.0, 1e-3, n), rng.uniform(2.0, 8.0, n), -1.0)
            return (fn, np.zeros(n), np.full(n, 2.0), -1.0)
        if fn == "mapstd":
            return (fn, rng.uniform(0.0, 0.05, n), rng.uniform(5.0, 20.0, n), 0.0 if rng.random() < 0.5 else 0.1)
        return (fn, None, None, 0.0)

    pin = [proc(fn, n_in, True) for fn in input_funcs]
    pout = [proc(fn, outputs, False) for fn in output_funcs]
    thr = [threshold] * outputs if np.isscalar(threshold) else list(threshold)
    return write_config(fs, fft_len, win_len, overlap, freq_range, time_range, thr, scaling, layers, pin, pout)
