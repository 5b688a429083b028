"""syldet-b200: B200-native syllable detection behind the reference's SyllableDetector / TrackDetector API.

Python host mirror of the Swift classes of gardner-lab/syllable-detector-swift, forwarding to the C-ABI of
libsyldet_cuda.so (include/syldet.h).  All compute happens in the CUDA library; there is NO CPU fallback and importing
this package raises if the library has not been built (run `python __graft_entry__.py` or
`python syllable-detector-swift_b200/build.py`).

    reference                                             here
    SyllableDetectorConfig(fromTextFile:)                 SyllableDetectorConfig(path) / .from_text(str)
    SyllableDetector(config:) .appendAudioData            SyllableDetector(config).append_audio_data(samples)
      .processNewValue() .lastOutputs .lastDetected         .process_new_value() .last_outputs .last_detected
      .seenSyllable()                                       .seen_syllable()
    TrackDetector(track:config:channel:) .process()       TrackDetector(samples, config, channel).process()
      .debounceTime / .debounceFrames                       .debounce_time / .debounce_frames
    main.swift loop over tracks                           BatchDetector(config).run(pcm[ch, n], ...)
    Processor.receiveAudioFrom (live, many channels)      StreamGroup(config, n_channels, max_buffer).submit(bufs)
    ResamplerLinear(fromRate:toRate:) .resampleVector     ResamplerLinear(from_rate, to_rate).resample_vector(x)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsyldet_cuda.so")

SCALING = ("linear", "log", "db")
TRANSFER = ("TanSig", "LogSig", "PureLin", "SatLin")
PROCESSING = ("mapminmax", "mapstd", "l2normalize", "normalize", "normalizestd")
LAYOUT_PLANAR, LAYOUT_INTERLEAVED = 0, 1
DETECT_ANY_OUTPUT, DETECT_FIRST_OUTPUT = 0, 1
PCM_F32, PCM_S16, PCM_S24 = 0, 1, 2
RESAMPLE_LINEAR, RESAMPLE_POLYPHASE = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FUSED, KERNEL_TENSOR = 0, 1, 2, 3
KERNEL_TENSOR_TF32 = 4   # tensor kernel with all three DFT products in TF32 (amplitude-invariant; see include/syldet.h)
KERNEL_WIDE = 5          # two-layer networks with a wide hidden layer on a hop-4 STFT: 3xTF32 tcgen05 contraction (kernels_wide.cu)
KERNEL_NAMES = {1: "generic", 2: "fused", 3: "tensor", 4: "tensor_tf32", 5: "wide"}

_STATUS = {1: "unableToOpenPath", 2: "missingValue", 3: "invalidValue", 4: "mismatchedLength", 5: "invalidConfiguration",
           6: "badArgument", 7: "cuda", 8: "outOfMemory", 9: "bufferOverflow", 10: "unsupported"}


class SyldetError(RuntimeError):
    """Raised for any non-zero syldet_status. `.status` is the code, `.kind` its name, `.key` the offending config key."""

    def __init__(self, status, message, key=""):
        super().__init__("%s: %s" % (_STATUS.get(status, status), message))
        self.status = status
        self.kind = _STATUS.get(status, str(status))
        self.key = key


class ParseError(SyldetError):
    """SyllableDetectorConfig.ParseError (SyllableDetectorConfig.swift:50-55); kind is the Swift case name."""


class Event(C.Structure):
    _fields_ = [("channel", C.c_int32), ("reserved", C.c_int32), ("sample", C.c_int64)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsyldet_cuda.so is not built (%s); there is no CPU fallback. Run `python __graft_entry__.py`." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, cp, i32, i64, dbl = C.c_void_p, C.c_char_p, C.c_int, C.c_int64, C.c_double
    pvp = C.POINTER(vp)
    sig = {
        "syldet_last_error": (cp, []), "syldet_version": (cp, []), "syldet_device_count": (i32, []),
        "syldet_config_error_key": (cp, []),
        "syldet_config_load_text": (i32, [cp, pvp]), "syldet_config_parse_text": (i32, [cp, C.c_size_t, pvp]),
        "syldet_config_free": (None, [vp]),
        "syldet_config_sampling_rate": (dbl, [vp]), "syldet_config_fourier_length": (i32, [vp]),
        "syldet_config_window_length": (i32, [vp]), "syldet_config_window_overlap": (i32, [vp]),
        "syldet_config_time_range": (i32, [vp]), "syldet_config_scaling": (i32, [vp]),
        "syldet_config_freq_range": (i32, [vp, C.POINTER(dbl), C.POINTER(dbl)]),
        "syldet_config_threshold_count": (i32, [vp]), "syldet_config_thresholds": (i32, [vp, vp, i32]),
        "syldet_config_net_inputs": (i32, [vp]), "syldet_config_net_outputs": (i32, [vp]),
        "syldet_config_layer_count": (i32, [vp]),
        "syldet_config_layer_info": (i32, [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "syldet_config_layer_weights": (i32, [vp, i32, vp, C.c_size_t]),
        "syldet_config_layer_biases": (i32, [vp, i32, vp, C.c_size_t]),
        "syldet_config_input_processing_count": (i32, [vp]), "syldet_config_output_processing_count": (i32, [vp]),
        "syldet_config_processing": (i32, [vp, i32, i32, C.POINTER(i32), C.POINTER(C.c_float), vp, vp, C.c_size_t]),
        "syldet_config_validate": (i32, [vp]),
        "syldet_config_freq_index_range": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
        "syldet_config_gap": (i32, [vp]), "syldet_config_hop": (i32, [vp]),
        "syldet_config_first_output_sample": (i64, [vp]), "syldet_config_num_columns": (i64, [vp, i64]),
        "syldet_config_num_evals": (i64, [vp, i64]), "syldet_config_debounce_frames": (i64, [vp, dbl]),
        "syldet_batch_create": (i32, [vp, i32, pvp]), "syldet_batch_destroy": (None, [vp]),
        "syldet_batch_set_slice_evals": (i32, [vp, i64]),
        "syldet_batch_set_kernel": (i32, [vp, i32]), "syldet_batch_active_kernel": (i32, [vp]),
        "syldet_batch_run_host": (i32, [vp, vp, i32, i32, i64, i64, i32, i64, i32, vp, pvp]),
        "syldet_batch_simulate_host": (i32, [vp, vp, i32, i32, i64, i64, i32, i32, vp]),
        "syldet_batch_launch_device": (i32, [vp, vp, i32, i64, i64, i32, i32, vp, vp]),
        "syldet_batch_collect": (i32, [vp, i64, pvp]), "syldet_batch_launch_count": (i64, [vp]), "syldet_plan_tensor_unit_tiles": (i32, [i64, i32, i32, i32]),
        "syldet_batch_last_detection_count": (i32, [vp, C.POINTER(i64)]),
        "syldet_batch_range_fallbacks": (i64, [vp]),
        "syldet_batch_wide_phase_ms": (i32, [vp, C.POINTER(dbl), C.POINTER(dbl)]),
        "syldet_batch_spectra_host": (i32, [vp, vp, i32, i32, i64, i64, i32, vp, C.POINTER(i64)]),
        "syldet_events_count": (i64, [vp]), "syldet_events_outputs_per_event": (i32, [vp]),
        "syldet_events_data": (C.POINTER(Event), [vp]), "syldet_events_outputs": (C.POINTER(C.c_float), [vp]),
        "syldet_events_free": (None, [vp]),
        "syldet_events_copy_columns": (None, [vp, vp, vp, vp]), "syldet_events_copy_compact": (i32, [vp, C.c_uint32, vp]),
        "syldet_detector_create": (i32, [vp, i32, pvp]), "syldet_detector_destroy": (None, [vp]),
        "syldet_detector_append": (i32, [vp, vp, i64]), "syldet_detector_process_new_value": (i32, [vp]),
        "syldet_detector_last_outputs": (i32, [vp, vp, i32]), "syldet_detector_last_detected": (i32, [vp]),
        "syldet_detector_seen_syllable": (i32, [vp]),
        "syldet_stream_create": (i32, [vp, i32, i32, i32, pvp]), "syldet_stream_destroy": (None, [vp]),
        "syldet_stream_create_resampled": (i32, [vp, i32, i32, i32, dbl, pvp]), "syldet_stream_resampling": (i32, [vp]),
        "syldet_stream_submit": (i32, [vp, vp, i32, vp, vp, vp]), "syldet_stream_launch_count": (i64, [vp]), "syldet_stream_fast_tick_count": (i64, [vp]), "syldet_stream_resident_tick_count": (i64, [vp]),
        "syldet_stream_read_levels": (i32, [vp, vp, vp]), "syldet_stream_set_pulse": (i32, [vp, dbl, dbl]),
        "syldet_stream_render_pulses": (i32, [vp, vp, i32]),
        "syldet_resampler_linear_create": (i32, [dbl, dbl, pvp]), "syldet_resampler_destroy": (None, [vp]),
        "syldet_resampler_process": (i32, [vp, vp, i64, vp, i64, C.POINTER(i64)]),
        "syldet_resampler_max_output": (i64, [vp, i64]),
        "syldet_resample_output_length": (i64, [i32, i64, dbl, dbl]),
        "syldet_resample_host": (i32, [i32, vp, i32, i64, i64, dbl, dbl, vp, i64, C.POINTER(i64), i32]),
        "syldet_resample_device": (i32, [i32, vp, i32, i64, i64, dbl, dbl, vp, i64, C.POINTER(i64), vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export what include/syldet.h declares
        fn.restype = res
        fn.argtypes = args
    L._syldet_signatures = sig
    return L


lib = _load()
EXPORTED_SYMBOLS = tuple(lib._syldet_signatures.keys())


def _check(status):
    if status != 0:
        msg = lib.syldet_last_error().decode(errors="replace")
        key = lib.syldet_config_error_key().decode(errors="replace")
        raise (ParseError if 1 <= status <= 4 else SyldetError)(status, msg, key)


def device_count():
    return lib.syldet_device_count()


def version():
    return lib.syldet_version().decode()


class SyllableDetectorConfig:
    """struct SyllableDetectorConfig + init(fromTextFile:) (Common/SyllableDetectorConfig.swift:11-45,170-277)."""

    def __init__(self, path=None, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            _check(lib.syldet_config_load_text(os.fsencode(path), C.byref(self._h)))

    @classmethod
    def from_text(cls, text):
        b = text.encode() if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        _check(lib.syldet_config_parse_text(b, len(b), C.byref(h)))
        return cls(_handle=h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib.syldet_config_free(self._h)
            self._h = None

    # -- parsed fields (Swift property names in snake case)
    sampling_rate = property(lambda s: lib.syldet_config_sampling_rate(s._h))
    fourier_length = property(lambda s: lib.syldet_config_fourier_length(s._h))
    window_length = property(lambda s: lib.syldet_config_window_length(s._h))
    window_overlap = property(lambda s: lib.syldet_config_window_overlap(s._h))
    time_range = property(lambda s: lib.syldet_config_time_range(s._h))
    spectrogram_scaling = property(lambda s: SCALING[lib.syldet_config_scaling(s._h)])
    net_inputs = property(lambda s: lib.syldet_config_net_inputs(s._h))
    net_outputs = property(lambda s: lib.syldet_config_net_outputs(s._h))
    layer_count = property(lambda s: lib.syldet_config_layer_count(s._h))

    @property
    def freq_range(self):
        lo, hi = C.c_double(), C.c_double()
        _check(lib.syldet_config_freq_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    @property
    def thresholds(self):
        n = lib.syldet_config_threshold_count(self._h)
        out = np.zeros(n, dtype=np.float64)
        _check(lib.syldet_config_thresholds(self._h, out.ctypes.data, n))
        return out

    def layer(self, i):
        """-> (weights [outputs, inputs] float32, biases float32, transfer function name)"""
        a, b, t = C.c_int(), C.c_int(), C.c_int()
        _check(lib.syldet_config_layer_info(self._h, i, C.byref(a), C.byref(b), C.byref(t)))
        w = np.zeros((b.value, a.value), dtype=np.float32)
        bias = np.zeros(b.value, dtype=np.float32)
        _check(lib.syldet_config_layer_weights(self._h, i, w.ctypes.data, w.size))
        _check(lib.syldet_config_layer_biases(self._h, i, bias.ctypes.data, bias.size))
        return w, bias, TRANSFER[t.value]

    def _processing(self, which, count, n):
        out = []
        for i in range(count):
            fn, y = C.c_int(), C.c_float()
            _check(lib.syldet_config_processing(self._h, which, i, C.byref(fn), C.byref(y), None, None, 0))
            if fn.value <= 1:
                xo, g = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
                _check(lib.syldet_config_processing(self._h, which, i, None, None, xo.ctypes.data, g.ctypes.data, n))
                out.append((PROCESSING[fn.value], xo, g, y.value))
            else:
                out.append((PROCESSING[fn.value], None, None, 0.0))
        return out

    @property
    def input_processing(self):
        return self._processing(0, lib.syldet_config_input_processing_count(self._h), self.net_inputs)

    @property
    def output_processing(self):
        return self._processing(1, lib.syldet_config_output_processing_count(self._h), self.net_outputs)

    # -- derived geometry (SyllableDetector.init / CircularShortTimeFourierTransform.init / TrackDetector.init)
    def validate(self):
        _check(lib.syldet_config_validate(self._h))
        return self

    @property
    def freq_indices(self):
        a, b = C.c_int(), C.c_int()
        _check(lib.syldet_config_freq_index_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    gap = property(lambda s: lib.syldet_config_gap(s._h))
    hop = property(lambda s: lib.syldet_config_hop(s._h))
    first_output_sample = property(lambda s: lib.syldet_config_first_output_sample(s._h))

    def num_columns(self, n_samples):
        return lib.syldet_config_num_columns(self._h, n_samples)

    def num_evals(self, n_samples):
        return lib.syldet_config_num_evals(self._h, n_samples)

    def debounce_frames(self, seconds):
        return lib.syldet_config_debounce_frames(self._h, float(seconds))


class Events:
    """CLI rows `channel,sample,seconds,out0[,out1...]` (TrackDetector.swift:92-96)."""

    def __init__(self, handle, sampling_rate):
        n = lib.syldet_events_count(handle)
        o = lib.syldet_events_outputs_per_event(handle)
        if n:   # the library splits its row array (syldet_event = {int32 channel, int32 reserved, int64 sample}) into columns in one pass
            self.channel = np.empty(n, dtype=np.int32)
            self.sample = np.empty(n, dtype=np.int64)
            self.outputs = np.empty((n, o), dtype=np.float32)
            lib.syldet_events_copy_columns(handle, self.channel.ctypes.data, self.sample.ctypes.data, self.outputs.ctypes.data)
        else:
            self.channel = np.zeros(0, dtype=np.int32)
            self.sample = np.zeros(0, dtype=np.int64)
            self.outputs = np.zeros((0, o), dtype=np.float32)
        lib.syldet_events_free(handle)
        self._sampling_rate = sampling_rate
        self._seconds = None

    @property
    def seconds(self):
        """sample / samplingRate (TrackDetector.swift:88-89 for a contiguous 0-based track); computed on first use"""
        if self._seconds is None:
            self._seconds = self.sample / self._sampling_rate
        return self._seconds

    def __len__(self):
        return self.sample.size

    def csv_lines(self):
        return ["%d,%d,%r,%s" % (c, s, float(t), ",".join(repr(float(np.float32(v))) for v in o))
                for c, s, t, o in zip(self.channel, self.sample, self.seconds, self.outputs)]


class BatchDetector:
    """All channels of a recording at once: TrackDetector.process + the main.swift loop (main.swift:126-130)."""

    def __init__(self, config, device=0, kernel=KERNEL_AUTO):
        self.config = config
        self._h = C.c_void_p()
        _check(lib.syldet_batch_create(config._h, device, C.byref(self._h)))
        if kernel != KERNEL_AUTO:
            _check(lib.syldet_batch_set_kernel(self._h, kernel))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.syldet_batch_destroy(self._h)
            self._h = None

    @property
    def active_kernel(self):
        return lib.syldet_batch_active_kernel(self._h)

    @staticmethod
    def available_kernels(config, device=0):
        """Kernel selectors this configuration qualifies for, fastest first."""
        out = []
        for k in (KERNEL_TENSOR, KERNEL_FUSED, KERNEL_WIDE, KERNEL_GENERIC):
            try:
                BatchDetector(config, device, kernel=k)
                out.append(k)
            except SyldetError as e:
                if e.kind != "unsupported":
                    raise
        return out

    @property
    def launch_count(self):
        return lib.syldet_batch_launch_count(self._h)

    def set_slice_evals(self, evals):
        """Minimum evaluations (all channels together) per time slice of run()'s copy/detect/collect pipeline."""
        _check(lib.syldet_batch_set_slice_evals(self._h, int(evals)))

    def run(self, pcm, debounce_frames=0, detect_rule=DETECT_ANY_OUTPUT, want_outputs=False, layout=LAYOUT_PLANAR):
        """pcm: host array, planar [n_channels, n_samples] (or interleaved [n_samples, n_channels]); float32 or int16.
        -> Events (, outputs [n_channels, E, O] float32)"""
        a = np.asarray(pcm)
        if a.ndim == 1:
            a = a[None, :] if layout == LAYOUT_PLANAR else a[:, None]
        fmt = PCM_S16 if a.dtype == np.int16 else PCM_F32
        a = np.ascontiguousarray(a, dtype=np.int16 if fmt == PCM_S16 else np.float32)
        nch, n = (a.shape if layout == LAYOUT_PLANAR else a.shape[::-1])
        E = self.config.num_evals(n)
        outs = np.zeros((nch, E, self.config.net_outputs), dtype=np.float32) if want_outputs else None
        ev = C.c_void_p()
        _check(lib.syldet_batch_run_host(self._h, a.ctypes.data, fmt, nch, n, n, layout, int(debounce_frames), detect_rule,
                                         outs.ctypes.data if want_outputs else None, C.byref(ev)))
        events = Events(ev, self.config.sampling_rate)
        return (events, outs) if want_outputs else events

    def run_into(self, table, recording, pcm, debounce_frames=0, detect_rule=DETECT_ANY_OUTPUT, layout=LAYOUT_PLANAR):
        """run(), with the detections appended to a sharding.EventTable as compact gather rows of `recording` - written by the library
        straight from its event list into the table's page-locked memory (no column arrays in between). -> number of detections"""
        a = np.asarray(pcm)
        if a.ndim == 1:
            a = a[None, :] if layout == LAYOUT_PLANAR else a[:, None]
        fmt = PCM_S16 if a.dtype == np.int16 else PCM_F32
        a = np.ascontiguousarray(a, dtype=np.int16 if fmt == PCM_S16 else np.float32)
        nch, n = (a.shape if layout == LAYOUT_PLANAR else a.shape[::-1])
        if not (0 <= int(recording) < 65536) or nch > 65536:
            raise ValueError("compact event rows hold recordings and channels below 65 536")
        ev = C.c_void_p()
        _check(lib.syldet_batch_run_host(self._h, a.ctypes.data, fmt, nch, n, n, layout, int(debounce_frames), detect_rule, None, C.byref(ev)))
        try:
            m = int(lib.syldet_events_count(ev))
            if lib.syldet_events_outputs_per_event(ev) != table.dtype["out"].shape[0]:
                raise ValueError("the table's rows hold a different number of outputs")
            addr, _ = table.claim(m)
            table.commit(m, lib.syldet_events_copy_compact(ev, int(recording), addr) if m else 1)
        finally:
            lib.syldet_events_free(ev)
        return m

    def simulate(self, pcm, s16=False, layout=LAYOUT_PLANAR):
        """Simulator trace (ViewControllerSimulator.swift:251-254, 308-344): clamp(out0 / thr0, 0, 1) held per hop, one value per
        input sample. -> [n_channels, n_samples] float32 (or int16 as the upstream 16-bit writer stores it)."""
        a = np.asarray(pcm)
        if a.ndim == 1:
            a = a[None, :] if layout == LAYOUT_PLANAR else a[:, None]
        fmt = PCM_S16 if a.dtype == np.int16 else PCM_F32
        a = np.ascontiguousarray(a, dtype=np.int16 if fmt == PCM_S16 else np.float32)
        nch, n = (a.shape if layout == LAYOUT_PLANAR else a.shape[::-1])
        trace = np.zeros((nch, n), dtype=np.int16 if s16 else np.float32)
        _check(lib.syldet_batch_simulate_host(self._h, a.ctypes.data, fmt, nch, n, n, layout, PCM_S16 if s16 else PCM_F32,
                                              trace.ctypes.data))
        return trace

    @property
    def range_fallbacks(self):
        """1 once the tensor kernel of this handle switched to its all-TF32 variant (audio outside the fp16 window), else 0."""
        return lib.syldet_batch_range_fallbacks(self._h)

    def wide_phase_ms(self):
        """-> (stft_ms, contraction_ms): device time of the two kernels of the last KERNEL_WIDE launch."""
        a, b = C.c_double(), C.c_double()
        _check(lib.syldet_batch_wide_phase_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def spectra(self, pcm, layout=LAYOUT_PLANAR):
        """extractPower()[f0 ..< f1] (CSTFT.swift:280-337) of every column that feeds an evaluation, as the active kernel computes it
        on the detection path. -> float32 [n_channels, num_evals + time_range - 1, band]"""
        a = np.asarray(pcm)
        if a.ndim == 1:
            a = a[None, :] if layout == LAYOUT_PLANAR else a[:, None]
        fmt = PCM_S16 if a.dtype == np.int16 else PCM_F32
        a = np.ascontiguousarray(a, dtype=np.int16 if fmt == PCM_S16 else np.float32)
        nch, n = (a.shape if layout == LAYOUT_PLANAR else a.shape[::-1])
        cols = C.c_int64()
        _check(lib.syldet_batch_spectra_host(self._h, a.ctypes.data, fmt, nch, n, n, layout, None, C.byref(cols)))
        k0, k1 = self.config.freq_indices
        band = np.zeros((nch, cols.value, k1 - k0), dtype=np.float32)
        if cols.value:
            _check(lib.syldet_batch_spectra_host(self._h, a.ctypes.data, fmt, nch, n, n, layout, band.ctypes.data, C.byref(cols)))
        return band

    def launch_device(self, d_pcm_ptr, n_channels, n_samples, channel_stride, detect_rule=DETECT_ANY_OUTPUT, d_outputs_ptr=None,
                      stream=None, layout=LAYOUT_PLANAR):
        """Asynchronous launch over device-resident float32 PCM (raw device pointers, cudaStream_t handle)."""
        _check(lib.syldet_batch_launch_device(self._h, d_pcm_ptr, n_channels, n_samples, channel_stride, layout, detect_rule,
                                              d_outputs_ptr, stream))

    def collect(self, debounce_frames=0):
        ev = C.c_void_p()
        _check(lib.syldet_batch_collect(self._h, int(debounce_frames), C.byref(ev)))
        return Events(ev, self.config.sampling_rate)

    def last_detection_count(self):
        n = C.c_int64()
        _check(lib.syldet_batch_last_detection_count(self._h, C.byref(n)))
        return n.value


class SyllableDetector:
    """class SyllableDetector (Common/SyllableDetector.swift:13-231)."""

    def __init__(self, config, device=0):
        self.config = config
        self._h = C.c_void_p()
        _check(lib.syldet_detector_create(config._h, device, C.byref(self._h)))
        self._n_out = config.net_outputs

    def __del__(self):
        if getattr(self, "_h", None):
            lib.syldet_detector_destroy(self._h)
            self._h = None

    def append_audio_data(self, samples):
        a = np.ascontiguousarray(samples, dtype=np.float32)
        _check(lib.syldet_detector_append(self._h, a.ctypes.data, a.size))

    def process_new_value(self):
        r = lib.syldet_detector_process_new_value(self._h)
        if r < 0:
            _check(-r)
        return r == 1

    @property
    def last_outputs(self):
        out = np.zeros(self._n_out, dtype=np.float32)
        _check(lib.syldet_detector_last_outputs(self._h, out.ctypes.data, out.size))
        return out

    @property
    def last_detected(self):
        return bool(lib.syldet_detector_last_detected(self._h))

    def seen_syllable(self):
        r = lib.syldet_detector_seen_syllable(self._h)
        if r < 0:
            _check(-r)
        return r == 1


class TrackDetector:
    """class TrackDetector (SyllableDetectorCLI/TrackDetector.swift:12-105). `source` stands in for the AVAssetTrack:
    an iterable of float32 sample buffers (any segmentation) already at config.sampling_rate."""

    def __init__(self, source, config, channel=0, device=0):
        self.detector = SyllableDetector(config, device)
        self.channel = channel
        self.debounce_frames = 0
        self._source = iter(source)
        self._next_output = config.first_output_sample  # TrackDetector.swift:39-42
        self._hop = config.hop
        self._total = 0
        self._until = -1
        self.rows = []  # (channel, sample, seconds, outputs)

    @property
    def debounce_time(self):
        return self.debounce_frames / self.detector.config.sampling_rate

    @debounce_time.setter
    def debounce_time(self, seconds):
        self.debounce_frames = self.detector.config.debounce_frames(seconds)

    def process(self):
        """One sample buffer (TrackDetector.swift:45-105). Returns False when the source is exhausted."""
        try:
            buf = next(self._source)
        except StopIteration:
            return False
        buf = np.ascontiguousarray(buf, dtype=np.float32)
        if buf.size == 0:
            return True
        cfg = self.detector.config
        thr = cfg.thresholds
        self.detector.append_audio_data(buf)
        while self.detector.process_new_value():
            cur = self._next_output
            self._next_output += self._hop
            outs = self.detector.last_outputs
            if any(float(d) >= thr[i] for i, d in enumerate(outs)) and self._until < cur:
                if cur - self._total >= buf.size:
                    raise SyldetError(5, "Unexpected sample number.")
                self.rows.append((self.channel, cur, cur / cfg.sampling_rate, outs.copy()))
                self._until = cur + self.debounce_frames
        self._total += buf.size
        return True


class StreamGroup:
    """Many live channels, one small buffer per channel per tick (SyllableDetector/Processor.swift:102-149)."""

    def __init__(self, config, n_channels, max_buffer=32, device=0, input_rate=None):
        """input_rate: sampling rate of the submitted buffers when it is not the configuration's (the audio device's rate,
        ViewControllerProcessor.swift:247-250): every buffer then goes through ResamplerLinear inside the tick kernel."""
        self.config = config
        self.n_channels = n_channels
        self._h = C.c_void_p()
        if input_rate is None:
            _check(lib.syldet_stream_create(config._h, n_channels, max_buffer, device, C.byref(self._h)))
        else:
            _check(lib.syldet_stream_create_resampled(config._h, n_channels, max_buffer, device, float(input_rate), C.byref(self._h)))
        self._seen = np.zeros(n_channels, dtype=np.uint8)
        self._n_new = np.zeros(n_channels, dtype=np.int32)
        self.last_outputs = np.zeros((n_channels, config.net_outputs), dtype=np.float32)
        self._ptrs = (C.c_void_p * n_channels)()

    def __del__(self):
        if getattr(self, "_h", None):
            lib.syldet_stream_destroy(self._h)
            self._h = None

    @property
    def launch_count(self):
        return lib.syldet_stream_launch_count(self._h)

    @property
    def fast_tick_count(self):
        """launches of the latency-shaped tick kernel among `launch_count`"""
        return lib.syldet_stream_fast_tick_count(self._h)

    @property
    def resident_tick_count(self):
        """ticks served by the resident tick kernel (SYLDET_STREAM_RESIDENT=1 at creation; no launch per tick)"""
        return lib.syldet_stream_resident_tick_count(self._h)

    @property
    def resampling(self):
        return bool(lib.syldet_stream_resampling(self._h))

    def read_levels(self):
        """-> (input_rms[n_channels], output_max[n_channels]) since the last call, NaN = upstream's nil
        (getInputForChannel / getOutputForChannel, Processor.swift:158-184); resets both."""
        a = np.zeros(self.n_channels, dtype=np.float64)
        b = np.zeros(self.n_channels, dtype=np.float64)
        _check(lib.syldet_stream_read_levels(self._h, a.ctypes.data, b.ctypes.data))
        return a, b

    def set_pulse(self, high_seconds=0.001, output_rate=44100.0):
        """ProcessorAudio.highDuration (Processor.swift:192): arm TTL pulses on detection."""
        _check(lib.syldet_stream_set_pulse(self._h, float(high_seconds), float(output_rate)))

    def render_pulses(self, n_frames):
        """The output device's render callback (AudioInterface.swift:13-40). -> float32 [n_channels, n_frames] of 1.0 / 0.0"""
        out = np.zeros((self.n_channels, n_frames), dtype=np.float32)
        ptrs = (C.c_void_p * self.n_channels)(*[out.ctypes.data + ch * n_frames * 4 for ch in range(self.n_channels)])
        _check(lib.syldet_stream_render_pulses(self._h, ptrs, n_frames))
        return out

    def submit(self, bufs):
        """bufs: float32 [n_channels, n]. -> (seen[n_channels] bool, n_new[n_channels])"""
        a = np.ascontiguousarray(bufs, dtype=np.float32)
        assert a.ndim == 2 and a.shape[0] == self.n_channels
        for ch in range(self.n_channels):
            self._ptrs[ch] = a.ctypes.data + ch * a.shape[1] * 4
        _check(lib.syldet_stream_submit(self._h, self._ptrs, a.shape[1], self._seen.ctypes.data, self._n_new.ctypes.data,
                                        self.last_outputs.ctypes.data))
        return self._seen.astype(bool), self._n_new.copy()


class ResamplerLinear:
    """class ResamplerLinear (Common/Resampler.swift:20-81)."""

    def __init__(self, from_rate, to_rate):
        self._h = C.c_void_p()
        _check(lib.syldet_resampler_linear_create(float(from_rate), float(to_rate), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.syldet_resampler_destroy(self._h)
            self._h = None

    def resample_vector(self, data):
        a = np.ascontiguousarray(data, dtype=np.float32)
        cap = lib.syldet_resampler_max_output(self._h, a.size)
        out = np.zeros(cap, dtype=np.float32)
        n = C.c_int64()
        _check(lib.syldet_resampler_process(self._h, a.ctypes.data, a.size, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    resample_array = resample_vector  # Resampler.swift:72-76


def resample(x, rate_in, rate_out, mode=RESAMPLE_POLYPHASE, device=0):
    """Whole-channel sample-rate conversion on the device. x: float32 [n] or [n_channels, n] (planar).
    RESAMPLE_POLYPHASE: rational Kaiser-windowed sinc converter (scipy.signal.resample_poly's algorithm) - what upstream gets from
    AVFoundation for files at another rate (Common/SyllableDetector.swift:19-23); RESAMPLE_LINEAR: one ResamplerLinear.resampleVector
    call per channel (Common/Resampler.swift:35-70)."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    one = a.ndim == 1
    if one:
        a = a[None, :]
    nch, n = a.shape
    n_out = lib.syldet_resample_output_length(mode, n, float(rate_in), float(rate_out))
    if n_out < 0:
        raise SyldetError(10, "the polyphase converter needs integral sampling rates")
    out = np.zeros((nch, max(n_out, 0)), dtype=np.float32)
    got = C.c_int64()
    _check(lib.syldet_resample_host(mode, a.ctypes.data, nch, n, n, float(rate_in), float(rate_out), out.ctypes.data, max(n_out, 1), C.byref(got), device))
    return out[0] if one else out
