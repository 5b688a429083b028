"""CPU-side tests of the C-ABI library: it loads, exports every symbol include/syldet.h declares, parses the
reference's text format exactly like the oracle's independent parser, and refuses to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, SAMPLE_TXT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "syldet.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(syldet_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(sd):
    names = _declared_symbols()
    assert len(names) > 50
    raw = ctypes.CDLL(sd.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing
    # and the Python mirror binds all of them
    assert set(names) == set(sd.EXPORTED_SYMBOLS)


def test_python_constants_mirror_the_header_enums(sd):
    """Every SYLDET_* enumerator the header defines with an explicit value has the same value in the Python mirror (the
    mirror drops the SYLDET_ prefix)."""
    text = open(os.path.join(ROOT, "include", "syldet.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    pairs = re.findall(r"\bSYLDET_([A-Z0-9_]+)\s*=\s*(-?\d+)", text)
    assert ("KERNEL_TENSOR_TF32", "4") in pairs and len(pairs) > 15
    checked = 0
    for name, value in pairs:
        if hasattr(sd, name):
            assert getattr(sd, name) == int(value), name
            checked += 1
    assert checked >= 10
    for name in ("KERNEL_AUTO", "KERNEL_GENERIC", "KERNEL_FUSED", "KERNEL_TENSOR", "KERNEL_TENSOR_TF32", "KERNEL_WIDE"):
        assert hasattr(sd, name), name


def test_sample_txt_fields(sd):
    c = sd.SyllableDetectorConfig(SAMPLE_TXT).validate()
    assert (c.sampling_rate, c.fourier_length, c.window_length, c.window_overlap, c.time_range) == (44100.0, 256, 256, 124, 10)
    assert c.freq_range == (2000.0, 7000.0) and c.spectrogram_scaling == "linear"
    assert c.freq_indices == (12, 41) and c.hop == 132 and c.gap == 0 and c.first_output_sample == 1444
    assert c.num_columns(2646000) == 20044 and c.num_evals(2646000) == 20035
    assert c.net_inputs == 290 and c.net_outputs == 1 and c.layer_count == 2
    assert [p[0] for p in c.input_processing] == ["l2normalize", "mapminmax"]
    assert c.debounce_frames(0.05) == 2205


def _assert_same_config(c, o):
    assert c.sampling_rate == o.samplingRate
    assert (c.fourier_length, c.window_length, c.window_overlap, c.time_range) == (o.fourierLength, o.windowLength, o.windowOverlap, o.timeRange)
    assert np.array_equal(c.thresholds, o.thresholds)
    assert c.freq_indices == (o.k0, o.k1) and c.hop == o.stride and c.gap == o.gap
    assert c.layer_count == o.layers
    for i in range(c.layer_count):
        w, b, tf = c.layer(i)
        assert np.array_equal(w.ravel().astype(np.float64), o.array("layer%d.weights" % i))  # bit-exact float32 parse
        assert np.array_equal(b.astype(np.float64), o.array("layer%d.biases" % i))
    for k, (fn, xo, g, y) in enumerate(c.input_processing):
        assert o.scalar("processInputs%d.function" % k) == ("mapminmax", "mapstd", "l2normalize", "normalize", "normalizestd").index(fn)
        if xo is not None:
            assert np.array_equal(xo.astype(np.float64), o.array("processInputs%d.xOffsets" % k))
            assert np.array_equal(g.astype(np.float64), o.array("processInputs%d.gains" % k))
            assert y == o.scalar("processInputs%d.y" % k)
    for k, (fn, xo, g, y) in enumerate(c.output_processing):
        assert np.array_equal(xo.astype(np.float64), o.array("processOutputs%d.xOffsets" % k))
        assert np.array_equal(g.astype(np.float64), o.array("processOutputs%d.gains" % k))


def test_parser_matches_oracle_parser(sd, oracle_mod, golden, sample_text):
    _assert_same_config(sd.SyllableDetectorConfig(SAMPLE_TXT).validate(), oracle_mod.Oracle(SAMPLE_TXT))
    for name, g in golden.items():
        _assert_same_config(sd.SyllableDetectorConfig.from_text(g["config"]).validate(), oracle_mod.Oracle(text=g["config"]))


def test_parser_matches_twin_parser(sd, sample_text):
    from oracle.twin64 import Twin64
    t = Twin64(SAMPLE_TXT)
    c = sd.SyllableDetectorConfig(SAMPLE_TXT)
    w, b, _ = c.layer(0)
    assert np.array_equal(w.astype(np.float64), t.layers[0][0]) and np.array_equal(b.astype(np.float64), t.layers[0][1])


MUTATIONS = [
    # (edit, expected ParseError case, key)   -- SyllableDetectorConfig.swift:50-55, 170-277
    (lambda s: s.replace("samplingRate = 44100.0\n", ""), "missingValue", "samplingRate"),
    (lambda s: s.replace("fourierLength = 256", "fourierLength = 250"), "invalidValue", "fourierLength"),
    (lambda s: s.replace("fourierLength = 256", "fourierLength = 256.0"), "invalidValue", "fourierLength"),
    (lambda s: s.replace("freqRange = 2000.0, 7000.0", "freqRange = 2000.0"), "mismatchedLength", "freqRange"),
    (lambda s: s.replace("freqRange = 2000.0, 7000.0", "freqRange = 2000.0, x"), "invalidValue", "freqRange"),
    (lambda s: s.replace("timeRange = 10", "timeRange = ten"), "invalidValue", "timeRange"),
    (lambda s: s.replace("threshold = ", "thresholdX = "), "missingValue", "threshold"),
    (lambda s: s.replace("scaling = linear", "scaling = Linear"), "invalidValue", "scaling"),
    (lambda s: s.replace("layer0.transferFunction = TanSig", "layer0.transferFunction = tansig"), "invalidValue", "layer0.transferFunction"),
    (lambda s: s.replace("layer1.biases = ", "layer1.biases = 0.5, "), "mismatchedLength", "layer1.biases"),
    (lambda s: s.replace("layer0.inputs = 290", "layer0.inputs = 291"), "mismatchedLength", "layer0.weights"),
    (lambda s: s.replace("processInputs0.function = l2normalize", "processInputs0.function = l3normalize"), "invalidValue", "processInputs0.function"),
    (lambda s: s.replace("processOutputs0.function = mapminmax", "processOutputs0.function = l2normalize"), "invalidValue", "processOutputs0.function"),
    (lambda s: s.replace("processInputs1.yMin = -1", ""), "missingValue", "processInputs1.yMin"),
    (lambda s: s.replace("processOutputsCount = 1", ""), "missingValue", "processOutputsCount"),
]


@pytest.mark.parametrize("idx", range(len(MUTATIONS)))
def test_parse_errors_match_reference_cases(sd, oracle_mod, sample_text, idx):
    edit, kind, key = MUTATIONS[idx]
    text = edit(sample_text)
    assert text != sample_text
    with pytest.raises(sd.ParseError) as ei:
        sd.SyllableDetectorConfig.from_text(text)
    assert (ei.value.kind, ei.value.key) == (kind, key)
    with pytest.raises(oracle_mod.OracleError) as eo:
        oracle_mod.Oracle(text=text)
    assert eo.value.code == ei.value.status and eo.value.key == key


def test_parser_quirks(sd, sample_text):
    C = sd.SyllableDetectorConfig
    # '#' lines are ignored only because they hold no '='; a second '=' in a value drops the line (:183-189)
    assert C.from_text("# note\n" + sample_text).fourier_length == 256
    with pytest.raises(sd.ParseError) as e:
        C.from_text(sample_text.replace("timeRange = 10", "timeRange = 10 = 10"))
    assert e.value.kind == "missingValue" and e.value.key == "timeRange"
    # "a == b" still splits into two pieces (empty pieces are dropped)
    assert C.from_text(sample_text.replace("timeRange = 10", "timeRange == 10")).time_range == 10
    # later duplicates win; CRLF line ends are trimmed; windowLength defaults to fourierLength (:204-209)
    assert C.from_text(sample_text + "timeRange = 7\n").time_range == 7
    assert C.from_text(sample_text.replace("\n", "\r\n")).time_range == 10
    assert C.from_text(sample_text.replace("windowLength = 256\n", "")).window_length == 256
    # `thresholds` wins over the legacy key, and falls back to it when unparsable (:223-229)
    assert list(C.from_text(sample_text + "thresholds = 0.25\n").thresholds) == [0.25]
    assert abs(C.from_text(sample_text + "thresholds = abc\n").thresholds[0] - 0.442442442442442) < 1e-15
    with pytest.raises(sd.ParseError) as e:
        C(os.path.join(ROOT, "does-not-exist.txt"))
    assert e.value.kind == "unableToOpenPath"


def test_invariants_are_status_codes_not_aborts(sd, sample_text):
    C = sd.SyllableDetectorConfig
    for edit in (lambda s: s.replace("windowOverlap = 124", "windowOverlap = 256"),      # CSTFT.swift:76
                 lambda s: s.replace("windowLength = 256", "windowLength = 300"),        # CSTFT.swift:86
                 lambda s: s.replace("freqRange = 2000.0, 7000.0", "freqRange = 7000.0, 2000.0"),  # SyllableDetector.swift:46
                 lambda s: s.replace("timeRange = 10", "timeRange = 9"),                 # SyllableDetector.swift:52
                 lambda s: s.replace("threshold = 0.442442442442442", "threshold = 0.4, 0.5")):  # SyllableDetector.swift:58
        c = C.from_text(edit(sample_text))
        with pytest.raises(sd.SyldetError) as e:
            c.validate()
        assert e.value.kind == "invalidConfiguration"


def test_negative_overlap_is_a_gap(sd, cw):
    c = sd.SyllableDetectorConfig.from_text(cw.random_config(seed=1, fft_len=512, win_len=400, overlap=-20, freq_range=(1000, 8000), time_range=6)).validate()
    assert c.gap == 20 and c.hop == 420 and c.first_output_sample == 400 + 420 * 5 + 20  # TrackDetector.swift:39-42
    assert c.num_columns(419) == 0 and c.num_columns(420) == 1


def test_compute_without_gpu_fails_loudly(sd):
    if sd.device_count() > 0:
        pytest.skip("a B200 is visible")
    c = sd.SyllableDetectorConfig(SAMPLE_TXT)
    for make in (lambda: sd.BatchDetector(c), lambda: sd.SyllableDetector(c), lambda: sd.StreamGroup(c, 4), lambda: sd.ResamplerLinear(48000, 44100)):
        with pytest.raises(sd.SyldetError) as e:
            make()
        assert e.value.kind == "cuda" and "no CPU fallback" in str(e.value)


def test_config_writer_round_trip(sd, cw):
    rng = np.random.default_rng(0)
    w0, b0 = rng.standard_normal((3, 58)), rng.standard_normal(3)
    w1, b1 = rng.standard_normal((2, 3)), rng.standard_normal(2)
    text = cw.write_config(44100.0, 256, 200, 100, (2000.0, 7000.0), 2, [0.1, 0.2], "db", [(w0, b0, "LogSig"), (w1, b1, "SatLin")],
                           [("normalize", None, None, 0), ("mapstd", rng.random(58), rng.random(58) + 1, 0.25)],
                           [("mapminmax", np.zeros(2), np.full(2, 2.0), -1.0)])
    c = sd.SyllableDetectorConfig.from_text(text).validate()
    w, b, tf = c.layer(0)
    assert tf == "LogSig" and np.array_equal(w, w0.astype(np.float32)) and np.array_equal(b, b0.astype(np.float32))
    assert c.layer(1)[2] == "SatLin" and c.spectrogram_scaling == "db" and list(c.thresholds) == [0.1, 0.2]
    assert [p[0] for p in c.input_processing] == ["normalize", "mapstd"] and c.input_processing[1][3] == 0.25


def test_tensor_unit_length_planner(sd):
    """launch_tc_range's choice of the unit length (host logic, no device): even, longer than four warm-ups, and never more tiles on the
    busiest of the round-robin CTAs than the former fixed 32 tiles per unit."""
    import math
    lib = sd.lib
    tf = 63   # frames per tile (64 hop rows)

    def busiest(tpu, evals, nch, sms, T):
        chunk = tpu * tf - (T - 1)
        return math.ceil(nch * math.ceil(max(evals, 1) / chunk) / sms) * (tpu + 1)

    assert lib.syldet_plan_tensor_unit_tiles(1202717, 8, 148, 10) == 74          # 1 h x 8 ch: 2 072 units = 14 on every CTA
    for evals, nch, sms, T in [(1202717, 8, 148, 10), (75169, 8, 148, 10), (20035, 1, 148, 10), (300, 2, 148, 10), (1202717, 1, 148, 10),
                               (5_000_000, 64, 148, 10), (20035, 3, 132, 30), (100000, 5, 148, 120), (0, 1, 148, 10)]:
        tpu = lib.syldet_plan_tensor_unit_tiles(evals, nch, sms, T)
        assert tpu >= 4 and tpu % 2 == 0 and tpu * tf > 4 * (T - 1)
        if 32 * tf > 4 * (T - 1):
            assert busiest(tpu, evals, nch, sms, T) <= busiest(32, evals, nch, sms, T)
    assert lib.syldet_plan_tensor_unit_tiles(1000, 1, 148, 3000) * tf > 4 * 2999   # very long window: the warm-up bound still holds
    assert lib.syldet_plan_tensor_unit_tiles(-1, 1, 148, 10) == 0


def test_resample_output_lengths(sd):
    """n_out of the two converters: Int(Float(n) / Float(rate_in / rate_out)) (Resampler.swift:32,40) and ceil(n up / down)."""
    assert sd.lib.syldet_resample_output_length(sd.RESAMPLE_LINEAR, 32, 48000.0, 44100.0) == int(np.float32(32) / np.float32(48000.0 / 44100.0))
    assert sd.lib.syldet_resample_output_length(sd.RESAMPLE_POLYPHASE, 48000, 48000.0, 44100.0) == 44100
    assert sd.lib.syldet_resample_output_length(sd.RESAMPLE_POLYPHASE, 100003, 96000.0, 44100.0) == -(-100003 * 147 // 320)
    assert sd.lib.syldet_resample_output_length(sd.RESAMPLE_POLYPHASE, 10, 44100.5, 44100.0) == -1
    assert sd.lib.syldet_resample_output_length(sd.RESAMPLE_LINEAR, 0, 48000.0, 44100.0) == 0
