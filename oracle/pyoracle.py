"""ctypes binding for oracle/liboracle.so (float32 restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c).  PARITY UNPINNED: the reference has no golden vectors
and cannot be built on Linux; the restatement is anchored by analytic KATs and a float64 twin.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile oracle.c -> liboracle.so with gcc (recipe = oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.c")
    out = lib_path()
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return out


def build_fast():
    """Timing-only build (-O3 -march=native, oracle/Makefile target `fast`) for bench.py's CPU legs. Compiled on the box that runs
    it, into a per-host file under the system temp directory (a library built for another CPU must not travel)."""
    import hashlib
    import platform
    import tempfile
    try:
        with open("/proc/cpuinfo") as f:
            cpu = next((l for l in f if l.startswith("flags")), platform.processor())
    except OSError:
        cpu = platform.processor()
    tag = hashlib.sha1((cpu + open(os.path.join(_HERE, "oracle.c")).read()).encode()).hexdigest()[:12]
    out = os.path.join(tempfile.gettempdir(), "liboracle_fast_%s.so" % tag)
    if not os.path.exists(out):
        subprocess.check_call(["make", "-C", _HERE, "-s", "fast", "FAST_OUT=" + out])
    return out


_libs = {}


def _load(fast=False):
    if fast in _libs:
        return _libs[fast]
    if fast:
        path = build_fast()
    else:
        if not os.path.exists(lib_path()):
            build()
        path = lib_path()
    L = C.CDLL(path)
    vp, cp, ip, lg, db = C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.c_long, C.c_double
    L.orc_config_parse.argtypes = [cp, C.c_size_t, C.POINTER(vp), ip, cp, C.c_int]
    L.orc_config_load.argtypes = [cp, C.POINTER(vp), ip, cp, C.c_int]
    L.orc_config_free.argtypes = [vp]
    L.orc_cfg_scalar.argtypes = [vp, cp]
    L.orc_cfg_scalar.restype = db
    L.orc_cfg_array.argtypes = [vp, cp, vp, lg]
    L.orc_cfg_array.restype = lg
    for fn in ("orc_num_columns", "orc_num_evals", "orc_eval_sample"):
        getattr(L, fn).argtypes = [vp, lg]
        getattr(L, fn).restype = lg
    L.orc_make_window.argtypes = [C.c_int, C.c_int, vp]
    L.orc_frame_spectrum.argtypes = [vp, vp, vp, vp]
    L.orc_stft_band.argtypes = [vp, vp, lg, vp]
    L.orc_stft_band.restype = lg
    L.orc_net_apply.argtypes = [vp, vp, vp]
    L.orc_run.argtypes = [vp, vp, lg, vp, vp, vp, vp]
    L.orc_run.restype = lg
    L.orc_run_multi.argtypes = [vp, vp, C.c_int, lg, lg, C.c_int, vp, vp]
    L.orc_run_multi.restype = lg
    L.orc_max_threads.restype = C.c_int
    L.orc_debounce.argtypes = [vp, vp, lg, lg, vp]
    L.orc_debounce.restype = lg
    L.orc_debounce_frames.argtypes = [vp, db]
    L.orc_debounce_frames.restype = lg
    L.orc_resampler_new.argtypes = [db, db]
    L.orc_resampler_new.restype = vp
    L.orc_resampler_free.argtypes = [vp]
    L.orc_resampler_process.argtypes = [vp, vp, lg, vp, lg]
    L.orc_resampler_process.restype = lg
    L.orc_resampler_state.argtypes = [vp, vp, vp, vp]
    _libs[fast] = L
    return L


class OracleError(Exception):
    """code: 1 unableToOpenPath, 2 missingValue, 3 invalidValue, 4 mismatchedLength, 5 invariant (fatalError upstream)."""

    def __init__(self, code, key):
        super().__init__("oracle config error %d at %r" % (code, key))
        self.code = code
        self.key = key


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    def __init__(self, path=None, text=None, fast=False):
        """fast=True: the timing-only -O3 -march=native build (bench.py's CPU legs); parity checks use the default build."""
        L = _load(fast)
        h = C.c_void_p()
        code = C.c_int(0)
        key = C.create_string_buffer(256)
        if path is not None:
            L.orc_config_load(os.fsencode(path), C.byref(h), C.byref(code), key, 256)
        else:
            b = text.encode() if isinstance(text, str) else bytes(text)
            L.orc_config_parse(b, len(b), C.byref(h), C.byref(code), key, 256)
        if code.value:
            raise OracleError(code.value, key.value.decode(errors="replace"))
        self._h = h
        self._L = L
        for n in ("fourierLength", "windowLength", "windowOverlap", "timeRange", "scaling", "layers", "gap", "overlap",
                  "stride", "k0", "k1", "L", "I", "O", "nThresholds"):
            setattr(self, n, int(self.scalar(n)))
        self.samplingRate = self.scalar("samplingRate")
        self.thresholds = self.array("thresholds")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_config_free(self._h)
            self._h = None

    def scalar(self, name):
        return float(self._L.orc_cfg_scalar(self._h, name.encode()))

    def array(self, name):
        n = self._L.orc_cfg_array(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n, dtype=np.float64)
        self._L.orc_cfg_array(self._h, name.encode(), out.ctypes.data, n)
        return out

    def num_columns(self, n):
        return int(self._L.orc_num_columns(self._h, n))

    def num_evals(self, n):
        return int(self._L.orc_num_evals(self._h, n))

    def eval_sample(self, j):
        return int(self._L.orc_eval_sample(self._h, j))

    def frame_spectrum(self, frame):
        frame = _f32(frame)
        assert frame.size == self.windowLength
        out = np.zeros(self.fourierLength // 2, dtype=np.float32)
        scratch = np.zeros(2 * self.fourierLength, dtype=np.float32)
        self._L.orc_frame_spectrum(self._h, frame.ctypes.data, out.ctypes.data, scratch.ctypes.data)
        return out

    def stft_band(self, x):
        x = _f32(x)
        Cn = self.num_columns(x.size)
        out = np.zeros((Cn, self.L), dtype=np.float32)
        if Cn:
            self._L.orc_stft_band(self._h, x.ctypes.data, x.size, out.ctypes.data)
        return out

    def net_apply(self, v):
        v = _f32(v)
        assert v.size == self.I
        out = np.zeros(self.O, dtype=np.float32)
        self._L.orc_net_apply(self._h, v.ctypes.data, out.ctypes.data)
        return out

    def run(self, x, want_band=False):
        """-> outputs[E,O] float32, det_any[E] bool, det_first[E] bool (, band[C,L])"""
        x = _f32(x)
        E, Cn = self.num_evals(x.size), self.num_columns(x.size)
        outs = np.zeros((E, self.O), dtype=np.float32)
        da = np.zeros(E, dtype=np.uint8)
        df = np.zeros(E, dtype=np.uint8)
        band = np.zeros((Cn, self.L), dtype=np.float32) if want_band else None
        self._L.orc_run(self._h, x.ctypes.data, x.size, outs.ctypes.data, da.ctypes.data, df.ctypes.data,
                        band.ctypes.data if want_band else None)
        r = (outs, da.astype(bool), df.astype(bool))
        return r + (band,) if want_band else r

    def run_multi(self, x, n_threads=0, want_outputs=True):
        """x: [n_channels, n] planar. -> outputs[ch,E,O], det_any[ch,E]"""
        x = _f32(x)
        nch, n = x.shape
        E = self.num_evals(n)
        outs = np.zeros((nch, E, self.O), dtype=np.float32) if want_outputs else None
        da = np.zeros((nch, E), dtype=np.uint8)
        self._L.orc_run_multi(self._h, x.ctypes.data, nch, n, n, n_threads,
                              outs.ctypes.data if want_outputs else None, da.ctypes.data)
        return outs, da.astype(bool)

    def max_threads(self):
        return int(self._L.orc_max_threads())

    def debounce_frames(self, seconds):
        return int(self._L.orc_debounce_frames(self._h, float(seconds)))

    def debounce(self, det, debounce_frames=0):
        """-> indices j of evaluations that are emitted as events (TrackDetector.swift:80,99)."""
        det = np.ascontiguousarray(det, dtype=np.uint8)
        ev = np.zeros(det.size, dtype=np.int64)
        n = self._L.orc_debounce(self._h, det.ctypes.data, det.size, int(debounce_frames), ev.ctypes.data)
        return ev[:n].copy()

    def events(self, x, debounce_frames=0, rule="any"):
        """CLI rows for one channel: (sample[int64], seconds[float64], outputs[n,O])."""
        outs, da, df = self.run(x)
        j = self.debounce(da if rule == "any" else df, debounce_frames)
        s = np.array([self.eval_sample(int(k)) for k in j], dtype=np.int64)
        return s, s / self.samplingRate, outs[j]


def simulator_trace(out0, thr0, first, hop, n_samples, s16=False):
    """Restates the simulator's output track (SyllableDetector/ViewControllerSimulator.swift:251-254 initial silent count,
    :326-344 value = clamp(lastOutputs[0] / Float(thresholds[0]), 0, 1) repeated for windowLength - windowOverlap samples).
    Samples past the last evaluation's hop are 0 here (upstream leaves stale buffer contents). Plain loop: small cases only."""
    tr = np.zeros(n_samples, dtype=np.float32)
    thr = np.float32(thr0)
    pos = first
    for v in np.asarray(out0, dtype=np.float32):
        if pos >= n_samples:
            break
        v = np.float32(v) / thr
        if v > 1.0:
            v = np.float32(1.0)
        elif v < 0.0:
            v = np.float32(0.0)
        tr[pos:min(pos + hop, n_samples)] = v
        pos += hop
    if not s16:
        return tr
    q = np.where(np.isnan(tr), np.float32(0.0), np.rint(tr * np.float32(32768.0)))
    return np.minimum(q, 32767.0).astype(np.int16)


def make_window(kind, n):
    """kind: 0 none, 1 hamming, 2 hann, 3 blackman (vDSP N-denominator forms)."""
    w = np.zeros(n, dtype=np.float32)
    _load().orc_make_window(kind, n, w.ctypes.data)
    return w


class Resampler:
    """ResamplerLinear restatement (Common/Resampler.swift:20-70), stateful across buffers."""

    def __init__(self, rate_in, rate_out):
        self._L = _load()
        self._h = self._L.orc_resampler_new(float(rate_in), float(rate_out))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_resampler_free(self._h)
            self._h = None

    def process(self, x):
        x = _f32(x)
        step, _, _ = self.state()
        cap = int(x.size / step) + 8
        out = np.zeros(cap, dtype=np.float32)
        n = self._L.orc_resampler_process(self._h, x.ctypes.data, x.size, out.ctypes.data, cap)
        assert n >= 0
        return out[:n].copy()

    def state(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self._L.orc_resampler_state(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value
