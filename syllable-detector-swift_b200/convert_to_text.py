"""Python counterpart of the reference's Matlab exporter `convert_to_text.m` plus a linter for the text format.

    convert_to_text(out_path, mat_path_or_dict, prepend_input_processing=())      # convert_to_text.m:1-214
    lint_config_text(text) -> [warnings]                                          # the parser's silent cases

The exporter reads a `.mat` file through scipy (structs only: a Matlab `network` object has to be saved as a struct
first, `net = struct(net)` on the Matlab side, which keeps the fields used here) or takes the equivalent dict:

    samplerate, fft_size, [win_size], fft_time_shift, freq_range, time_window_steps, trigger_thresholds, scaling,
    net.input/.output {processFcns, processSettings[{xoffset, gain, ymin|ymean}]}, net.layers[{netInputFcn, transferFcn}],
    net.IW, net.LW, net.b

and applies the same defaults, checks and `%.15g` formatting, so the text it writes is byte-identical to what the
Matlab function writes for the same numbers (convert_to_text.m:61-74 header, :118-182 processing, :184-212 layers).
Pure host-side tooling: nothing here runs on the detection path.
"""
import warnings

import numpy as np

from .config_writer import INPUT_FUNCS, TRANSFER, write_config

_TRANSFER_NAMES = {"tansig": "TanSig", "logsig": "LogSig", "purelin": "PureLin", "satlin": "SatLin"}  # convert_to_text.m:189-199


class ConvertError(ValueError):
    """The `error(...)` calls of convert_to_text.m."""


def _get(obj, name, default=None):
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


def _has(obj, name):
    return (name in obj) if isinstance(obj, dict) else hasattr(obj, name)


def _cells(v):
    """A Matlab cell array / struct array as a Python list (scipy squeezes 1-element cells to the element itself)."""
    if v is None:
        return []
    if isinstance(v, (list, tuple)):
        return list(v)
    if isinstance(v, np.ndarray) and v.dtype == object:
        return list(v.ravel())
    if isinstance(v, np.ndarray) and v.size == 0:
        return []
    return [v]


def _cell2d(v, rows, cols):
    """IW / LW / b cell arrays as a rows x cols list of arrays (empty cells -> size-0 arrays)."""
    if isinstance(v, np.ndarray) and v.dtype == object:
        a = v.reshape(rows, cols)
        return [[np.atleast_1d(np.asarray(a[i, j], dtype=np.float64)) for j in range(cols)] for i in range(rows)]
    if isinstance(v, (list, tuple)):
        flat = [np.atleast_1d(np.asarray(x, dtype=np.float64)) for row in v for x in (row if isinstance(row, (list, tuple)) else [row])]
        if len(flat) != rows * cols:
            raise ConvertError("cell array of unexpected size")
        return [flat[i * cols:(i + 1) * cols] for i in range(rows)]
    if rows * cols == 1:
        return [[np.atleast_1d(np.asarray(v, dtype=np.float64))]]
    raise ConvertError("cell array of unexpected size")


def load_mat(path):
    """scipy.io.loadmat with Matlab structs as attribute objects and singleton dimensions squeezed."""
    from scipy.io import loadmat
    return {k: v for k, v in loadmat(path, struct_as_record=False, squeeze_me=True).items() if not k.startswith("__")}


def _processing(put, pre=()):
    """convert_processing_functions (convert_to_text.m:118-182) -> [(function, xOffsets, gains, y)]"""
    fcns = [str(f) for f in _cells(_get(put, "processFcns"))]
    settings = _cells(_get(put, "processSettings"))
    items = [(str(name), None, None, 0.0) for name in pre]       # :136-142: prepended functions are written by name only
    if len(fcns) + len(items) == 0:
        warnings.warn("Zero processing functions no longer results in linear normalization of input vectors.")  # :129-131
    for fn, st in zip(fcns, settings):
        if fn == "mapminmax":
            items.append((fn, np.atleast_1d(_get(st, "xoffset")), np.atleast_1d(_get(st, "gain")), float(_get(st, "ymin"))))
        elif fn == "mapstd":
            items.append((fn, np.atleast_1d(_get(st, "xoffset")), np.atleast_1d(_get(st, "gain")), float(_get(st, "ymean"))))
        else:
            raise ConvertError("Invalid processing function: %s." % fn)  # :167-168
    return items


def convert_to_text(fn, mat, prepend_input_processing=()):
    """convert_to_text(fn, mat, 'prepend_input_processing', {...}). `mat` is a path or an already loaded dict. Returns the text
    (and writes it to `fn` unless fn is None)."""
    f = load_mat(mat) if isinstance(mat, (str, bytes)) or hasattr(mat, "__fspath__") else dict(mat)
    if isinstance(prepend_input_processing, str):
        prepend_input_processing = (prepend_input_processing,)   # :14-16
    fft_size = int(f["fft_size"])
    win_size = int(f.get("win_size", fft_size))                  # :33-35
    if fft_size <= 0 or fft_size & (fft_size - 1):
        raise ConvertError("Only FFT sizes that are a power of two are supported.")       # :40-42
    if win_size > fft_size:
        raise ConvertError("The window size must be less than or equal to the FFT size.")  # :45-47
    if 256 > fft_size:
        warnings.warn("The spectrogram defaults to using an FFT size of 256. As a result, the provided FFT size will be ignored.")
        fft_size = 256                                           # :50-53
    net = f["net"]
    freq = np.atleast_1d(np.asarray(f["freq_range"], dtype=np.float64)).ravel()
    layers_meta = _cells(_get(net, "layers"))
    n = len(layers_meta)
    IW = _cell2d(_get(net, "IW"), n, 1)
    LW = _cell2d(_get(net, "LW"), n, n)
    B = _cell2d(_get(net, "b"), n, 1)
    layers = []
    for i, meta in enumerate(layers_meta):
        if any(LW[i][j].size for j in range(n) if j != i - 1):
            raise ConvertError("Networks with only connections between consecutive layers supported.")  # :93-95
        if i == 0:
            w = IW[0][0]
        else:
            w = LW[i][i - 1]
            if IW[i][0].size:
                raise ConvertError("Found unexpected input weights for layer 1.")        # :102-104
        b = B[i][0].ravel()
        if str(_get(meta, "netInputFcn", "netsum")) != "netsum":
            raise ConvertError("Invalid input function: %s. Expected netsum." % _get(meta, "netInputFcn"))  # :185-187
        tf = _TRANSFER_NAMES.get(str(_get(meta, "transferFcn")))
        if tf is None:
            raise ConvertError("Invalid transfer function: %s." % _get(meta, "transferFcn"))  # :198-199
        w = np.asarray(w, dtype=np.float64)
        w = w.reshape(b.size, -1) if w.ndim != 2 else w          # a 1 x n or n x 1 matrix arrives squeezed
        layers.append((w, b, tf))
    text = write_config(float(f["samplerate"]), fft_size, win_size, fft_size - int(f["fft_time_shift"]), (freq[0], freq[-1]),
                        int(f["time_window_steps"]), np.atleast_1d(np.asarray(f["trigger_thresholds"], dtype=np.float64)).ravel(),
                        str(f["scaling"]), layers, _processing(_get(net, "input"), prepend_input_processing),
                        _processing(_get(net, "output")))
    if fn is not None:
        with open(fn, "w") as fh:
            fh.write(text)
    return text


def lint_config_text(text):
    """Reports what SyllableDetectorConfig.init(fromTextFile:) accepts silently (SyllableDetectorConfig.swift:183-189 and the
    helpers at :57-168): lines dropped because they do not split into exactly two pieces at '=', duplicate keys (last one wins),
    keys no parser rule reads, the legacy `threshold` key, counts that disagree with the lists that follow."""
    out, seen = [], {}
    for no, line in enumerate(text.splitlines(), 1):
        pieces = [p for p in line.split("=") if p != ""]          # Swift split drops empty pieces
        if len(pieces) != 2:
            if line.strip() and not line.lstrip().startswith("#"):
                out.append("line %d is ignored: it does not have the form key = value (%d '=' separated pieces)" % (no, len(pieces)))
            elif line.lstrip().startswith("#") and len(pieces) == 2:
                pass
            continue
        key, val = pieces[0].strip(), pieces[1].strip()
        if line.lstrip().startswith("#"):
            out.append("line %d starts with '#' but contains one '=', so it is parsed as the key %r" % (no, key))
        if key in seen:
            out.append("line %d: key %r repeats line %d; the later value wins" % (no, key, seen[key][0]))
        seen[key] = (no, val)

    def known(k):
        import re
        if k in ("samplingRate", "fourierLength", "windowLength", "windowOverlap", "freqRange", "timeRange", "thresholds", "threshold",
                 "scaling", "layers", "processInputsCount", "processOutputsCount"):
            return True
        return bool(re.fullmatch(r"layer\d+\.(inputs|outputs|weights|biases|transferFunction)", k) or
                    re.fullmatch(r"process(Inputs|Outputs)\d+\.(function|xOffsets|gains|yMin|yMean)", k))

    for k, (no, _) in seen.items():
        if not known(k):
            out.append("line %d: key %r is not read by the parser" % (no, k))
    if "threshold" in seen and "thresholds" not in seen:
        out.append("line %d: legacy key 'threshold' (still accepted, SyllableDetectorConfig.swift:223-229)" % seen["threshold"][0])
    if "windowLength" not in seen:
        out.append("windowLength is absent: it defaults to fourierLength (SyllableDetectorConfig.swift:204-209)")

    def count(key):
        try:
            return int(seen[key][1])
        except (KeyError, ValueError):
            return None

    n_layers = count("layers")
    if n_layers is not None:
        extra = [k for k in seen if k.startswith("layer") and "." in k and k[5:k.index(".")].isdigit() and int(k[5:k.index(".")]) >= n_layers]
        if extra:
            out.append("layers = %d but keys of later layers exist and are ignored: %s" % (n_layers, ", ".join(sorted(extra)[:4])))
    for which in ("Inputs", "Outputs"):
        c = count("process%sCount" % which)
        pre = "process%s" % which
        idx = sorted({int(k[len(pre):k.index(".")]) for k in seen if k.startswith(pre) and "." in k and k[len(pre):k.index(".")].isdigit()})
        if c is None and idx:
            out.append("%sCount is absent: the %d %s entries are ignored" % (pre, len(idx), pre))
        elif c is not None and idx and idx[-1] >= c:
            out.append("%sCount = %d but entries up to index %d exist; the extra ones are ignored" % (pre, c, idx[-1]))
        for i in idx:
            fn = seen.get("%s%d.function" % (pre, i), (0, ""))[1]
            allowed = INPUT_FUNCS if which == "Inputs" else ("mapminmax", "mapstd")
            if fn and fn not in allowed:
                out.append("%s%d.function = %s is not accepted for process%s (allowed: %s)" % (pre, i, fn, which, ", ".join(allowed)))
    for k, (no, v) in seen.items():
        if k.endswith(".transferFunction") and v not in TRANSFER:
            out.append("line %d: transfer function %r is not one of %s" % (no, v, ", ".join(TRANSFER)))
    return out
