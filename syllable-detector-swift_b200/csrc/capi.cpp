// extern "C" surface of libsyldet_cuda.so (include/syldet.h). Thin forwarding only.
#include <cstring>
#include <new>

#include "engine.hpp"
#include <algorithm>

#include "stream.hpp"

using namespace syldet;

namespace {
template <typename F>
syldet_status guarded(F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return set_error(SYLDET_ERR_NOMEM, "out of host memory");
    } catch (const std::exception &e) {
        return set_error(SYLDET_ERR_ARG, std::string("unexpected exception: ") + e.what());
    }
}
const Config *valid_or_null(const syldet_config *cfg) { return cfg ? &cfg->c : nullptr; }
}  // namespace

extern "C" {

const char *syldet_last_error(void) { return last_error_message().c_str(); }
const char *syldet_config_error_key(void) { return last_error_key().c_str(); }
const char *syldet_version(void) { return "syldet-b200 0.1.0 (sm_100a)"; }
int syldet_device_count(void) { return usable_device_count(); }

syldet_status syldet_config_load_text(const char *path, syldet_config **out) {
    if (!path || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    *out = nullptr;
    return guarded([&] {
        auto *cfg = new syldet_config();
        syldet_status st = load_config_file(path, cfg->c);
        if (st != SYLDET_OK) { delete cfg; return st; }
        *out = cfg;
        return SYLDET_OK;
    });
}

syldet_status syldet_config_parse_text(const char *text, size_t len, syldet_config **out) {
    if (!text || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    *out = nullptr;
    return guarded([&] {
        auto *cfg = new syldet_config();
        syldet_status st = parse_config_text(text, len, cfg->c);
        if (st != SYLDET_OK) { delete cfg; return st; }
        *out = cfg;
        return SYLDET_OK;
    });
}

void syldet_config_free(syldet_config *cfg) { delete cfg; }

double syldet_config_sampling_rate(const syldet_config *cfg) { return cfg->c.sampling_rate; }
int syldet_config_fourier_length(const syldet_config *cfg) { return cfg->c.fourier_length; }
int syldet_config_window_length(const syldet_config *cfg) { return cfg->c.window_length; }
int syldet_config_window_overlap(const syldet_config *cfg) { return cfg->c.window_overlap; }
int syldet_config_time_range(const syldet_config *cfg) { return cfg->c.time_range; }
int syldet_config_scaling(const syldet_config *cfg) { return cfg->c.scaling; }
syldet_status syldet_config_freq_range(const syldet_config *cfg, double *lo, double *hi) {
    if (!cfg || !lo || !hi) return set_error(SYLDET_ERR_ARG, "null argument");
    *lo = cfg->c.freq_lo;
    *hi = cfg->c.freq_hi;
    return SYLDET_OK;
}
int syldet_config_threshold_count(const syldet_config *cfg) { return (int)cfg->c.thresholds.size(); }
syldet_status syldet_config_thresholds(const syldet_config *cfg, double *out, int cap) {
    if (!cfg || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    if (cap < (int)cfg->c.thresholds.size()) return set_error(SYLDET_ERR_ARG, "capacity too small");
    std::memcpy(out, cfg->c.thresholds.data(), cfg->c.thresholds.size() * sizeof(double));
    return SYLDET_OK;
}
int syldet_config_net_inputs(const syldet_config *cfg) { return cfg->c.layers.empty() ? 0 : cfg->c.layers.front().inputs; }
int syldet_config_net_outputs(const syldet_config *cfg) { return cfg->c.layers.empty() ? 0 : cfg->c.layers.back().outputs; }
int syldet_config_layer_count(const syldet_config *cfg) { return (int)cfg->c.layers.size(); }
syldet_status syldet_config_layer_info(const syldet_config *cfg, int layer, int *inputs, int *outputs, int *transfer) {
    if (!cfg || layer < 0 || layer >= (int)cfg->c.layers.size()) return set_error(SYLDET_ERR_ARG, "layer index out of range");
    if (inputs) *inputs = cfg->c.layers[layer].inputs;
    if (outputs) *outputs = cfg->c.layers[layer].outputs;
    if (transfer) *transfer = cfg->c.layers[layer].transfer;
    return SYLDET_OK;
}
syldet_status syldet_config_layer_weights(const syldet_config *cfg, int layer, float *w, size_t cap) {
    if (!cfg || !w || layer < 0 || layer >= (int)cfg->c.layers.size()) return set_error(SYLDET_ERR_ARG, "layer index out of range");
    const auto &v = cfg->c.layers[layer].weights;
    if (cap < v.size()) return set_error(SYLDET_ERR_ARG, "capacity too small");
    std::memcpy(w, v.data(), v.size() * sizeof(float));
    return SYLDET_OK;
}
syldet_status syldet_config_layer_biases(const syldet_config *cfg, int layer, float *b, size_t cap) {
    if (!cfg || !b || layer < 0 || layer >= (int)cfg->c.layers.size()) return set_error(SYLDET_ERR_ARG, "layer index out of range");
    const auto &v = cfg->c.layers[layer].biases;
    if (cap < v.size()) return set_error(SYLDET_ERR_ARG, "capacity too small");
    std::memcpy(b, v.data(), v.size() * sizeof(float));
    return SYLDET_OK;
}
int syldet_config_input_processing_count(const syldet_config *cfg) { return (int)cfg->c.input_processing.size(); }
int syldet_config_output_processing_count(const syldet_config *cfg) { return (int)cfg->c.output_processing.size(); }
syldet_status syldet_config_processing(const syldet_config *cfg, int which, int index, int *function, float *y, float *xoff,
                                       float *gain, size_t cap) {
    if (!cfg || (which != 0 && which != 1)) return set_error(SYLDET_ERR_ARG, "bad chain selector");
    const auto &chain = which == 0 ? cfg->c.input_processing : cfg->c.output_processing;
    if (index < 0 || index >= (int)chain.size()) return set_error(SYLDET_ERR_ARG, "processing index out of range");
    const Processing &p = chain[index];
    if (function) *function = p.function;
    if (y) *y = p.y;
    if (xoff) {
        if (cap < p.x_offsets.size()) return set_error(SYLDET_ERR_ARG, "capacity too small");
        std::memcpy(xoff, p.x_offsets.data(), p.x_offsets.size() * sizeof(float));
    }
    if (gain) {
        if (cap < p.gains.size()) return set_error(SYLDET_ERR_ARG, "capacity too small");
        std::memcpy(gain, p.gains.data(), p.gains.size() * sizeof(float));
    }
    return SYLDET_OK;
}

syldet_status syldet_config_validate(const syldet_config *cfg) {
    if (!cfg) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] { return validate_config(const_cast<syldet_config *>(cfg)->c); });
}
static bool ensure_valid(const syldet_config *cfg) {
    return cfg && (cfg->c.valid || validate_config(const_cast<syldet_config *>(cfg)->c) == SYLDET_OK);
}
syldet_status syldet_config_freq_index_range(const syldet_config *cfg, int *start, int *end) {
    if (!cfg || !start || !end) return set_error(SYLDET_ERR_ARG, "null argument");
    if (!frequency_index_range(cfg->c.fourier_length, cfg->c.freq_lo, cfg->c.freq_hi, cfg->c.sampling_rate, *start, *end))
        return set_error(SYLDET_ERR_CONFIG, "The frequency range is invalid.", "freqRange");
    return SYLDET_OK;
}
int syldet_config_gap(const syldet_config *cfg) { return ensure_valid(cfg) ? cfg->c.gap : -1; }
int syldet_config_hop(const syldet_config *cfg) { return ensure_valid(cfg) ? cfg->c.hop : -1; }
int64_t syldet_config_first_output_sample(const syldet_config *cfg) { return ensure_valid(cfg) ? cfg->c.first_output_sample() : -1; }
int64_t syldet_config_num_columns(const syldet_config *cfg, int64_t n) { return ensure_valid(cfg) ? cfg->c.num_columns(n) : -1; }
int64_t syldet_config_num_evals(const syldet_config *cfg, int64_t n) { return ensure_valid(cfg) ? cfg->c.num_evals(n) : -1; }
int64_t syldet_config_debounce_frames(const syldet_config *cfg, double seconds) {
    return (int64_t)(seconds * cfg->c.sampling_rate);  // Int(newValue * samplingRate), TrackDetector.swift:23-25
}

// ---- batch ---------------------------------------------------------------------------------------------------------
syldet_status syldet_batch_create(const syldet_config *cfg, int device, syldet_batch **out) {
    if (!valid_or_null(cfg) || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    *out = nullptr;
    return guarded([&] {
        auto *b = new syldet_batch();
        syldet_status st = b->b.init(cfg->c, device);
        if (st != SYLDET_OK) { delete b; return st; }
        *out = b;
        return SYLDET_OK;
    });
}
void syldet_batch_destroy(syldet_batch *b) { delete b; }
syldet_status syldet_batch_set_kernel(syldet_batch *b, int kernel) {
    if (!b) return set_error(SYLDET_ERR_ARG, "null argument");
    return b->b.set_kernel(kernel);
}
int syldet_batch_active_kernel(const syldet_batch *b) { return b ? b->b.active_kernel() : -1; }
syldet_status syldet_batch_set_slice_evals(syldet_batch *b, int64_t evals) {
    if (!b || evals <= 0) return set_error(SYLDET_ERR_ARG, "slice size must be positive");
    b->b.set_slice_evals(evals);
    return SYLDET_OK;
}

syldet_status syldet_batch_run_host(syldet_batch *b, const void *pcm, int pcm_format, int n_channels, int64_t n_samples,
                                    int64_t channel_stride, int layout, int64_t debounce_frames, int detect_rule,
                                    float *all_outputs, syldet_events **events) {
    if (!b || !events) return set_error(SYLDET_ERR_ARG, "null argument");
    *events = nullptr;
    return guarded([&] {
        auto *ev = new syldet_events();
        syldet_status st = b->b.run_host(pcm, pcm_format, n_channels, n_samples, channel_stride, layout, debounce_frames,
                                         detect_rule, all_outputs, ev->e);
        if (st != SYLDET_OK) { delete ev; return st; }
        *events = ev;
        return SYLDET_OK;
    });
}
syldet_status syldet_batch_simulate_host(syldet_batch *b, const void *pcm, int pcm_format, int n_channels, int64_t n_samples,
                                         int64_t channel_stride, int layout, int trace_format, void *trace) {
    if (!b || !trace) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] {
        Events ev;
        return b->b.run_host(pcm, pcm_format, n_channels, n_samples, channel_stride, layout, 0, SYLDET_DETECT_FIRST_OUTPUT, nullptr, ev,
                             trace_format, trace);
    });
}
syldet_status syldet_batch_launch_device(syldet_batch *b, const float *d_pcm, int n_channels, int64_t n_samples,
                                         int64_t channel_stride, int layout, int detect_rule, float *d_all_outputs, void *stream) {
    if (!b) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] {
        return b->b.launch_device(d_pcm, n_channels, n_samples, channel_stride, layout, detect_rule, d_all_outputs,
                                  static_cast<cudaStream_t>(stream));
    });
}
syldet_status syldet_batch_collect(syldet_batch *b, int64_t debounce_frames, syldet_events **events) {
    if (!b || !events) return set_error(SYLDET_ERR_ARG, "null argument");
    *events = nullptr;
    return guarded([&] {
        auto *ev = new syldet_events();
        syldet_status st = b->b.collect(debounce_frames, ev->e);
        if (st != SYLDET_OK) { delete ev; return st; }
        *events = ev;
        return SYLDET_OK;
    });
}
int64_t syldet_batch_launch_count(const syldet_batch *b) { return b ? b->b.launch_count() : 0; }
int syldet_plan_tensor_unit_tiles(int64_t evals_per_channel, int n_channels, int sm_count, int time_range) {
    if (evals_per_channel < 0 || n_channels <= 0 || sm_count <= 0 || time_range <= 0) return 0;
    return tc_plan_unit_tiles(evals_per_channel, n_channels, sm_count, tc_tile_frames(), time_range - 1);
}
int64_t syldet_batch_range_fallbacks(const syldet_batch *b) { return b ? b->b.range_fallbacks() : 0; }
syldet_status syldet_batch_wide_phase_ms(syldet_batch *b, double *stft_ms, double *contraction_ms) {
    if (!b) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] { return b->b.wide_phase_ms(stft_ms, contraction_ms); });
}
syldet_status syldet_batch_spectra_host(syldet_batch *b, const void *pcm, int pcm_format, int n_channels, int64_t n_samples,
                                        int64_t channel_stride, int layout, float *band, int64_t *n_columns) {
    if (!b) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] { return b->b.spectra_host(pcm, pcm_format, n_channels, n_samples, channel_stride, layout, band, n_columns); });
}
syldet_status syldet_batch_last_detection_count(syldet_batch *b, int64_t *count) {
    if (!b || !count) return set_error(SYLDET_ERR_ARG, "null argument");
    return b->b.last_detection_count(count);
}

int64_t syldet_events_count(const syldet_events *ev) { return ev ? (int64_t)ev->e.rows.size() : 0; }
int syldet_events_outputs_per_event(const syldet_events *ev) { return ev ? ev->e.outputs_per_event : 0; }
const syldet_event *syldet_events_data(const syldet_events *ev) { return ev ? ev->e.rows.data() : nullptr; }
const float *syldet_events_outputs(const syldet_events *ev) { return ev ? ev->e.outputs.data() : nullptr; }
void syldet_events_copy_columns(const syldet_events *ev, int32_t *channel, int64_t *sample, float *outputs) {
    if (!ev) return;
    const size_t n = ev->e.rows.size();
    const syldet_event *r = ev->e.rows.data();
    if (channel && sample) {
        for (size_t i = 0; i < n; ++i) {
            channel[i] = r[i].channel;
            sample[i] = r[i].sample;
        }
    } else {
        if (channel) for (size_t i = 0; i < n; ++i) channel[i] = r[i].channel;
        if (sample) for (size_t i = 0; i < n; ++i) sample[i] = r[i].sample;
    }
    if (outputs && n) std::memcpy(outputs, ev->e.outputs.data(), ev->e.outputs.size() * sizeof(float));
}
int syldet_events_copy_compact(const syldet_events *ev, uint32_t recording, void *rows) {
    if (!ev || !rows) return 1;
    const size_t n = ev->e.rows.size();
    const int O = ev->e.outputs_per_event;
    const size_t item = 12 + 4 * (size_t)O;
    const syldet_event *r = ev->e.rows.data();
    const float *o = ev->e.outputs.data();
    unsigned char *dst = static_cast<unsigned char *>(rows);
    int ordered = 1;
    uint32_t prev_key = 0;
    int64_t prev_sample = 0;
    for (size_t i = 0; i < n; ++i, dst += item) {
        const uint32_t key = (recording << 16) | (uint32_t)(r[i].channel & 0xffff);
        const int64_t smp = r[i].sample;
        if (i && (key < prev_key || (key == prev_key && smp < prev_sample))) ordered = 0;
        prev_key = key;
        prev_sample = smp;
        std::memcpy(dst, &key, 4);
        std::memcpy(dst + 4, &smp, 8);
        std::memcpy(dst + 12, o + i * (size_t)O, 4 * (size_t)O);
    }
    return ordered;
}
void syldet_events_free(syldet_events *ev) { delete ev; }

// ---- batched resampling -------------------------------------------------------------------------------------------------
int64_t syldet_resample_output_length(int mode, int64_t n_in, double rate_in, double rate_out) {
    return resample_output_length(mode, n_in, rate_in, rate_out);
}
syldet_status syldet_resample_host(int mode, const float *in, int n_channels, int64_t n_in, int64_t in_stride, double rate_in, double rate_out,
                                   float *out, int64_t out_stride, int64_t *n_out, int device) {
    return guarded([&] { return resample_host(mode, in, n_channels, n_in, in_stride, rate_in, rate_out, out, out_stride, n_out, device); });
}
syldet_status syldet_resample_device(int mode, const float *d_in, int n_channels, int64_t n_in, int64_t in_stride, double rate_in,
                                     double rate_out, float *d_out, int64_t out_stride, int64_t *n_out, void *stream) {
    return guarded([&] {
        static thread_local DeviceBuffer *filter = new DeviceBuffer();   // the polyphase taps of this thread's last conversion
        return resample_device(mode, d_in, n_channels, n_in, in_stride, rate_in, rate_out, d_out, out_stride, n_out, *filter,
                               static_cast<cudaStream_t>(stream));
    });
}

// ---- single stream ---------------------------------------------------------------------------------------------------
syldet_status syldet_detector_create(const syldet_config *cfg, int device, syldet_detector **out) {
    if (!valid_or_null(cfg) || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    *out = nullptr;
    return guarded([&] {
        auto *d = new syldet_detector();
        syldet_status st = d->d.init(cfg->c, device);
        if (st != SYLDET_OK) { delete d; return st; }
        *out = d;
        return SYLDET_OK;
    });
}
void syldet_detector_destroy(syldet_detector *d) { delete d; }
syldet_status syldet_detector_append(syldet_detector *d, const float *samples, int64_t n) {
    if (!d) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] { return d->d.append(samples, n); });
}
int syldet_detector_process_new_value(syldet_detector *d) {
    if (!d) return -(int)set_error(SYLDET_ERR_ARG, "null argument");
    try {
        return d->d.process_new_value();
    } catch (const std::exception &e) {
        return -(int)set_error(SYLDET_ERR_NOMEM, e.what());
    }
}
syldet_status syldet_detector_last_outputs(const syldet_detector *d, float *out, int cap) {
    if (!d || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    const auto &v = d->d.last_outputs();
    if (cap < (int)v.size()) return set_error(SYLDET_ERR_ARG, "capacity too small");
    std::memcpy(out, v.data(), v.size() * sizeof(float));
    return SYLDET_OK;
}
int syldet_detector_last_detected(const syldet_detector *d) { return d && d->d.last_detected() ? 1 : 0; }
int syldet_detector_seen_syllable(syldet_detector *d) {
    if (!d) return -(int)set_error(SYLDET_ERR_ARG, "null argument");
    try {
        return d->d.seen_syllable();
    } catch (const std::exception &e) {
        return -(int)set_error(SYLDET_ERR_NOMEM, e.what());
    }
}

// ---- live group ------------------------------------------------------------------------------------------------------
syldet_status syldet_stream_create(const syldet_config *cfg, int n_channels, int max_buffer, int device, syldet_stream **out) {
    if (!valid_or_null(cfg) || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    *out = nullptr;
    return guarded([&] {
        auto *s = new syldet_stream();
        syldet_status st = s->g.init(cfg->c, n_channels, max_buffer, device);
        if (st != SYLDET_OK) { delete s; return st; }
        *out = s;
        return SYLDET_OK;
    });
}
syldet_status syldet_stream_create_resampled(const syldet_config *cfg, int n_channels, int max_buffer, int device, double input_rate,
                                            syldet_stream **out) {
    if (!valid_or_null(cfg) || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    *out = nullptr;
    return guarded([&] {
        auto *s = new syldet_stream();
        syldet_status st = s->g.init(cfg->c, n_channels, max_buffer, device, input_rate);
        if (st != SYLDET_OK) { delete s; return st; }
        *out = s;
        return SYLDET_OK;
    });
}
int syldet_stream_resampling(const syldet_stream *s) { return s && s->g.resampling() ? 1 : 0; }
void syldet_stream_destroy(syldet_stream *s) { delete s; }
syldet_status syldet_stream_submit(syldet_stream *s, const float *const *bufs, int n, uint8_t *seen, int32_t *n_new, float *last_out) {
    if (!s || !bufs) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] {
        const float *outs = nullptr;
        int64_t fresh = 0;
        syldet_status st = s->g.submit(bufs, n, &outs, &fresh);
        if (st != SYLDET_OK) return st;
        const Config &c = s->g.config();
        const int O = c.outputs, nch = s->g.n_channels();
        for (int ch = 0; ch < nch; ++ch) {
            bool any = false;
            for (int64_t j = 0; j < fresh; ++j)  // lastDetected for every new value (Processor.swift:136-144)
                if ((double)outs[((size_t)ch * fresh + j) * O] >= c.thresholds[0]) any = true;
            if (seen) seen[ch] = any ? 1 : 0;
            if (any && s->high_frames > 0) s->high_for[ch] = s->high_frames;  // createHighOutput (AudioInterface.swift:442-445)
            if (n_new) n_new[ch] = (int32_t)fresh;
            if (last_out && fresh > 0) std::memcpy(last_out + (size_t)ch * O, outs + ((size_t)ch * fresh + fresh - 1) * O, O * sizeof(float));
        }
        return SYLDET_OK;
    });
}
int64_t syldet_stream_launch_count(const syldet_stream *s) { return s ? s->g.launch_count() : 0; }
int64_t syldet_stream_fast_tick_count(const syldet_stream *s) { return s ? s->g.fast_tick_count() : 0; }
int64_t syldet_stream_resident_tick_count(const syldet_stream *s) { return s ? s->g.resident_tick_count() : 0; }
syldet_status syldet_stream_read_levels(syldet_stream *s, double *input_rms, double *output_max) {
    if (!s) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] { return s->g.read_levels(input_rms, output_max); });
}
syldet_status syldet_stream_set_pulse(syldet_stream *s, double high_seconds, double output_rate) {
    if (!s || !(high_seconds >= 0.0) || !(output_rate > 0.0)) return set_error(SYLDET_ERR_ARG, "bad pulse arguments");
    return guarded([&] {
        s->high_frames = (int64_t)(high_seconds * output_rate);  // Int(duration * outputFormat.mSampleRate)
        s->high_for.assign(s->g.n_channels(), 0);
        return SYLDET_OK;
    });
}
syldet_status syldet_stream_render_pulses(syldet_stream *s, float *const *out, int n_frames) {
    if (!s || !out || n_frames < 0) return set_error(SYLDET_ERR_ARG, "bad render arguments");
    if (s->high_for.empty()) return set_error(SYLDET_ERR_ARG, "syldet_stream_set_pulse has not been called");
    for (int ch = 0; ch < s->g.n_channels(); ++ch) {  // renderOutput (AudioInterface.swift:23-36)
        const int64_t high = s->high_for[ch];
        if (0 < high) s->high_for[ch] = high - std::min<int64_t>(high, n_frames);
        if (out[ch])
            for (int i = 0; i < n_frames; ++i) out[ch][i] = i < high ? 1.0f : 0.0f;
    }
    return SYLDET_OK;
}

// ---- resampler -------------------------------------------------------------------------------------------------------
syldet_status syldet_resampler_linear_create(double rate_in, double rate_out, syldet_resampler **out) {
    if (!out) return set_error(SYLDET_ERR_ARG, "null argument");
    *out = nullptr;
    return guarded([&] {
        auto *r = new syldet_resampler();
        syldet_status st = r->r.init(rate_in, rate_out, 0);
        if (st != SYLDET_OK) { delete r; return st; }
        *out = r;
        return SYLDET_OK;
    });
}
void syldet_resampler_destroy(syldet_resampler *r) { delete r; }
syldet_status syldet_resampler_process(syldet_resampler *r, const float *in, int64_t n_in, float *out, int64_t cap, int64_t *n_out) {
    if (!r || !out || !n_out) return set_error(SYLDET_ERR_ARG, "null argument");
    return guarded([&] { return r->r.process(in, n_in, out, cap, n_out); });
}
int64_t syldet_resampler_max_output(const syldet_resampler *r, int64_t n_in) { return r ? r->r.max_output(n_in) : 0; }

}  // extern "C"
