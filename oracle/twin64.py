"""float64 numpy twin of the hot path — TEST INFRASTRUCTURE ONLY (see oracle/oracle.c).

Same maths as SURVEY.md Appendix C, evaluated in float64 with numpy's FFT.  Used to bound the float32
error of both the C oracle and the CUDA path and to flag near-threshold evaluations.  It has its own
pure-Python parser so the C parsers can be cross-checked against an independent reading of
Common/SyllableDetectorConfig.swift:170-277.
"""
import ctypes
import math

import numpy as np

_libc = ctypes.CDLL(None)
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p)]


def _strtof(s):
    """decimal -> float32 in one rounding (Swift Float(String)); float(np.float32(float(s))) would round twice."""
    b = s.encode()
    end = ctypes.c_char_p()
    v = _libc.strtof(b, ctypes.byref(end))
    if not b or b[:1].isspace() or end.value not in (b"", None):
        raise ValueError(s)
    return float(v)


_WS = " \t\n\r\v\f"


def parse_kv(text):
    kv = {}
    for line in text.split("\n"):
        parts = [p for p in line.split("=") if p != ""]
        if len(parts) == 2:
            kv[parts[0].strip(_WS)] = parts[1].strip(_WS)
    return kv


class Twin64:
    def __init__(self, path=None, text=None):
        if text is None:
            with open(path, "rb") as f:
                text = f.read().decode("utf-8", errors="replace")
        kv = parse_kv(text)
        self.kv = kv

        def farr(key, n=None):
            vals = [_strtof(p.strip(_WS)) for p in kv[key].split(",") if p != ""]
            if n is not None:
                assert len(vals) == n, key
            return np.array(vals, dtype=np.float64)

        self.fs = float(kv["samplingRate"])
        self.N = int(kv["fourierLength"])
        self.W = int(kv["windowLength"]) if "windowLength" in kv else self.N
        ov = int(kv["windowOverlap"])
        self.gap, self.overlap = (-ov, 0) if ov < 0 else (0, ov)
        self.stride = self.gap + self.W - self.overlap
        lo, hi = [float(p.strip(_WS)) for p in kv["freqRange"].split(",") if p != ""]
        self.T = int(kv["timeRange"])
        key = "thresholds" if "thresholds" in kv else "threshold"
        self.thr = np.array([float(p.strip(_WS)) for p in kv[key].split(",") if p != ""])
        self.scaling = kv["scaling"]
        half = self.N // 2
        frm = self.N / self.fs
        self.k0 = int(math.ceil(frm * lo))
        self.k1 = min(int(math.floor(frm * hi)) + 1, half)
        self.L = self.k1 - self.k0
        self.layers = []
        for i in range(int(kv["layers"])):
            ni, no = int(kv["layer%d.inputs" % i]), int(kv["layer%d.outputs" % i])
            self.layers.append((farr("layer%d.weights" % i, ni * no).reshape(no, ni), farr("layer%d.biases" % i, no),
                                kv["layer%d.transferFunction" % i]))
        self.I = self.layers[0][0].shape[1]
        self.O = self.layers[-1][0].shape[0]

        def procs(prefix, n):
            out = []
            for i in range(int(kv[prefix + "Count"])):
                nm = "%s%d" % (prefix, i)
                fn = kv[nm + ".function"]
                if fn in ("mapminmax", "mapstd"):
                    y = _strtof(kv[nm + (".yMin" if fn == "mapminmax" else ".yMean")])
                    out.append((fn, farr(nm + ".xOffsets", n), farr(nm + ".gains", n), y))
                else:
                    out.append((fn, None, None, 0.0))
            return out

        self.ip = procs("processInputs", self.I)
        self.op = procs("processOutputs", self.O)
        n = np.arange(self.W)
        # float32 table like the oracle/GPU (the window is data, not arithmetic)
        self.window = (0.54 - 0.46 * np.cos(2 * np.pi * n / self.W)).astype(np.float32).astype(np.float64)

    def num_columns(self, n):
        need = self.gap + self.W
        return 0 if n < need else (n - need) // self.stride + 1

    def band(self, x):
        x = np.asarray(x, dtype=np.float64)
        C = self.num_columns(x.size)
        if C == 0:
            return np.zeros((0, self.L))
        idx = (np.arange(C) * self.stride + self.gap)[:, None] + np.arange(self.W)[None, :]
        fr = np.zeros((C, self.N))
        fr[:, :self.W] = x[idx] * self.window
        return np.abs(np.fft.rfft(fr, axis=1))[:, self.k0:self.k1]

    def net(self, v):
        """v: [E, I] scaled features -> [E, O]"""
        x = np.array(v, dtype=np.float64)
        with np.errstate(all="ignore"):
            for fn, xo, g, y in self.ip:
                if fn == "mapminmax":
                    x = (x - xo) * g + y
                elif fn == "mapstd":
                    x = (x - xo) * g + y
                elif fn == "l2normalize":
                    x = x / np.sqrt((x * x).sum(axis=1, keepdims=True))
                elif fn == "normalize":
                    mn, mx = x.min(axis=1, keepdims=True), x.max(axis=1, keepdims=True)
                    r = mx - mn
                    x = np.where(r == 0, -1.0, x * (2.0 / r) + (0 - mn - mx) / r)
                elif fn == "normalizestd":
                    x = (x - x.mean(axis=1, keepdims=True)) / x.std(axis=1, keepdims=True)
            for w, b, tf in self.layers:
                x = x @ w.T + b
                if tf == "TanSig":
                    x = np.tanh(x)
                elif tf == "LogSig":
                    x = 1.0 / (1.0 + np.exp(-x))
                elif tf == "SatLin":
                    x = np.clip(x, 0.0, 1.0)
            for fn, xo, g, y in self.op:
                x = (x - y) / g + xo
        return x

    def features(self, band):
        E = band.shape[0] - self.T + 1
        if E <= 0:
            return np.zeros((0, self.I))
        idx = np.arange(E)[:, None] + np.arange(self.T)[None, :]
        f = band[idx].reshape(E, self.I)
        with np.errstate(all="ignore"):
            if self.scaling == "db":
                f = 20.0 * np.log10(f)
            elif self.scaling == "log":
                f = np.log(f)
        return f

    def run(self, x):
        """-> outputs[E, O] float64"""
        return self.net(self.features(self.band(x)))

    def eval_sample(self, j):
        return self.gap + self.W + self.stride * (self.T - 1) + self.stride * np.asarray(j)
