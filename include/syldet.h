/*
 * syldet.h — C-ABI of libsyldet_cuda.so, the B200 (sm_100a) syllable-detection engine.
 *
 * The reference (gardner-lab/syllable-detector-swift) has no FFI boundary for this path: the Swift classes in
 * Common/ are compiled straight into the app and the CLI, and the only C boundary is the bridging header for
 * TPCircularBuffer (Common/Common-Bridging-Header.h:5).  This header therefore defines the boundary a Swift-on-Linux
 * shim binds through a module map (include/module.modulemap); each entry point names the reference interface it
 * replaces.  Paths are relative to the reference root.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns a syldet_status (0 = ok) unless it is a pure getter;
 *   - syldet_last_error() returns a thread-local, human-readable description of the last failure on this thread;
 *   - invariant violations that are fatalError() upstream are reported as SYLDET_ERR_CONFIG / SYLDET_ERR_OVERFLOW, the
 *     Swift shim re-raises them as fatalError to keep upstream behaviour;
 *   - audio is IEEE float32 at the configuration's sampling rate (upstream asks AVFoundation for exactly that:
 *     Common/SyllableDetector.swift:19-23);
 *   - there is NO CPU fallback: every compute entry point fails with SYLDET_ERR_CUDA when no sm_100 device is usable.
 *   - calls on one handle must be serialised by the caller; distinct handles are independent; a syldet_config is
 *     immutable and may be shared (stronger than upstream, where NeuralNet scratch is shared: NeuralNet.swift:241).
 */
#ifndef SYLDET_H
#define SYLDET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum syldet_status {
    SYLDET_OK = 0,
    SYLDET_ERR_OPEN = 1,        /* ParseError.unableToOpenPath  (Common/SyllableDetectorConfig.swift:51) */
    SYLDET_ERR_MISSING = 2,     /* ParseError.missingValue      (:52) */
    SYLDET_ERR_INVALID = 3,     /* ParseError.invalidValue      (:53) */
    SYLDET_ERR_MISMATCH = 4,    /* ParseError.mismatchedLength  (:54) */
    SYLDET_ERR_CONFIG = 5,      /* fatalError invariants: SyllableDetector.swift:46-60, CSTFT.swift:76-91, NeuralNet.swift:245-255,340-348 */
    SYLDET_ERR_ARG = 6,         /* bad argument to this API */
    SYLDET_ERR_CUDA = 7,        /* CUDA failure or no usable sm_100 device (there is no CPU fallback) */
    SYLDET_ERR_NOMEM = 8,
    SYLDET_ERR_OVERFLOW = 9,    /* "Insufficient space on buffer." fatalError (CSTFT.swift:199, SyllableDetector.swift:146) */
    SYLDET_ERR_UNSUPPORTED = 10
} syldet_status;

enum { SYLDET_SCALING_LINEAR = 0, SYLDET_SCALING_LOG = 1, SYLDET_SCALING_DB = 2 };         /* SyllableDetectorConfig.Scaling :13-30 */
enum { SYLDET_TF_TANSIG = 0, SYLDET_TF_LOGSIG = 1, SYLDET_TF_PURELIN = 2, SYLDET_TF_SATLIN = 3 }; /* NeuralNet.swift:189-228 */
enum { SYLDET_PROC_MAPMINMAX = 0, SYLDET_PROC_MAPSTD = 1, SYLDET_PROC_L2NORMALIZE = 2,
       SYLDET_PROC_NORMALIZE = 3, SYLDET_PROC_NORMALIZESTD = 4 };                          /* NeuralNet.swift:41-182 */
enum { SYLDET_LAYOUT_PLANAR = 0, SYLDET_LAYOUT_INTERLEAVED = 1 };
enum { SYLDET_DETECT_ANY_OUTPUT = 0,   /* TrackDetector.swift:71-77 (CLI rule)  */
       SYLDET_DETECT_FIRST_OUTPUT = 1  /* SyllableDetector.lastDetected :27-31 (live rule) */ };
enum { SYLDET_PCM_F32 = 0, SYLDET_PCM_S16 = 1 /* int16, x / 32768 */, SYLDET_PCM_S24 = 2 /* packed little-endian 24-bit, x / 2^23 */ };
enum { SYLDET_RESAMPLE_LINEAR = 0,    /* ResamplerLinear (Common/Resampler.swift:20-70), bit-faithful */
       SYLDET_RESAMPLE_POLYPHASE = 1  /* rational Kaiser-windowed sinc converter (the algorithm of scipy.signal.resample_poly) */ };
enum { SYLDET_KERNEL_AUTO = 0, SYLDET_KERNEL_GENERIC = 1, SYLDET_KERNEL_FUSED = 2, SYLDET_KERNEL_TENSOR = 3, SYLDET_KERNEL_TENSOR_TF32 = 4,
       SYLDET_KERNEL_WIDE = 5 /* wide hidden layer as a 3xTF32 tcgen05 contraction over a high-overlap STFT (hop 4) */ };

typedef struct syldet_config syldet_config;     /* SyllableDetectorConfig + NeuralNet                    */
typedef struct syldet_batch syldet_batch;       /* TrackDetector + main.swift loop, many channels at once */
typedef struct syldet_detector syldet_detector; /* SyllableDetector (one stream)                          */
typedef struct syldet_stream syldet_stream;     /* Processor.swift: many live channels, small buffers     */
typedef struct syldet_resampler syldet_resampler; /* ResamplerLinear                                       */
typedef struct syldet_events syldet_events;

/* One CSV row of the CLI: "channel,sample,seconds,out0[,out1...]" (TrackDetector.swift:92-96); seconds = sample / fs. */
typedef struct syldet_event {
    int32_t channel;
    int32_t reserved;
    int64_t sample; /* S_j = gap + W + stride*(T-1) + stride*j (TrackDetector.swift:39-42,67-68) */
} syldet_event;

const char *syldet_last_error(void);
const char *syldet_version(void);
/* Number of usable CUDA devices with compute capability 10.x; 0 when none. */
int syldet_device_count(void);

/* ---- configuration: SyllableDetectorConfig.init(fromTextFile:) (SyllableDetectorConfig.swift:170-277) ------------ */
syldet_status syldet_config_load_text(const char *path, syldet_config **out);
syldet_status syldet_config_parse_text(const char *text, size_t len, syldet_config **out);
void syldet_config_free(syldet_config *cfg);
/* name of the key the last parse error on this thread refers to (the String payload of ParseError) */
const char *syldet_config_error_key(void);

double syldet_config_sampling_rate(const syldet_config *cfg);
int syldet_config_fourier_length(const syldet_config *cfg);
int syldet_config_window_length(const syldet_config *cfg);
int syldet_config_window_overlap(const syldet_config *cfg); /* raw value; negative = gap (CSTFT.swift:66-73) */
int syldet_config_time_range(const syldet_config *cfg);
int syldet_config_scaling(const syldet_config *cfg);
syldet_status syldet_config_freq_range(const syldet_config *cfg, double *lo, double *hi);
int syldet_config_threshold_count(const syldet_config *cfg);
syldet_status syldet_config_thresholds(const syldet_config *cfg, double *out, int cap);
int syldet_config_net_inputs(const syldet_config *cfg);  /* NeuralNet.inputs  (NeuralNet.swift:235) */
int syldet_config_net_outputs(const syldet_config *cfg); /* NeuralNet.outputs (:236) */
int syldet_config_layer_count(const syldet_config *cfg);
syldet_status syldet_config_layer_info(const syldet_config *cfg, int layer, int *inputs, int *outputs, int *transfer);
syldet_status syldet_config_layer_weights(const syldet_config *cfg, int layer, float *w, size_t cap); /* row-major [out][in] */
syldet_status syldet_config_layer_biases(const syldet_config *cfg, int layer, float *b, size_t cap);
int syldet_config_input_processing_count(const syldet_config *cfg);
int syldet_config_output_processing_count(const syldet_config *cfg);
/* which: 0 = input chain, 1 = output chain. xoff/gain may be NULL; they receive `n` values for mapminmax/mapstd. */
syldet_status syldet_config_processing(const syldet_config *cfg, int which, int index, int *function, float *y,
                                       float *xoff, float *gain, size_t cap);

/* The checks SyllableDetector.init performs (SyllableDetector.swift:37-60, CSTFT.swift:61-129). Also run by every *_create. */
syldet_status syldet_config_validate(const syldet_config *cfg);
/* frequencyIndexRange (CSTFT.swift:166-191) of the configured band. */
syldet_status syldet_config_freq_index_range(const syldet_config *cfg, int *start, int *end);
int syldet_config_gap(const syldet_config *cfg);
int syldet_config_hop(const syldet_config *cfg);                               /* gap + W - overlap */
int64_t syldet_config_first_output_sample(const syldet_config *cfg);           /* TrackDetector.swift:39-42 */
int64_t syldet_config_num_columns(const syldet_config *cfg, int64_t n_samples);
int64_t syldet_config_num_evals(const syldet_config *cfg, int64_t n_samples);
/* Int(seconds * samplingRate) (TrackDetector.swift:23-25) */
int64_t syldet_config_debounce_frames(const syldet_config *cfg, double seconds);

/* ---- batch: what TrackDetector.process + main.swift do, for n_channels independent channels ---------------------- */
syldet_status syldet_batch_create(const syldet_config *cfg, int device, syldet_batch **out);
void syldet_batch_destroy(syldet_batch *b);
/* SYLDET_KERNEL_AUTO picks the fastest kernel the configuration qualifies for: TENSOR (tcgen05 band DFT + fused epilogue),
 * FUSED (SIMT FFT + fused epilogue), WIDE (two-layer networks with up to 1024 hidden units on a hop-4 STFT: the hidden layer as a
 * 3xTF32 tcgen05 contraction), else the GENERIC reference-order path.
 * TENSOR computes the two correction products of its 3xTF32 band DFT in fp16 for the reference's sample network shape (l2normalize
 * first). That pass is at float32 level inside an amplitude window (no sample beyond +-32752; norm of the band-magnitude window
 * >= 2^-8, i.e. audio rms >~ 2e-5; any 16-bit PCM input qualifies by construction). The kernel checks the window on every
 * evaluation; when audio falls outside it, syldet_batch_collect / _last_detection_count / _run_host transparently repeat the launch
 * with the all-TF32 variant and the handle keeps using that variant (syldet_batch_range_fallbacks counts the switch). Results are
 * therefore amplitude-invariant like the reference's float32 arithmetic; dense device outputs of syldet_batch_launch_device are final
 * once one of those calls has returned. TENSOR_TF32 selects the all-TF32 variant up front (~10 % slower); the environment variable
 * SYLDET_TC_TF32_CORR=1 does the same for every handle. */
syldet_status syldet_batch_set_kernel(syldet_batch *b, int kernel);
int syldet_batch_active_kernel(const syldet_batch *b);
/* syldet_batch_run_host cuts a recording into up to 16 time slices so that the PCIe copy of slice k+1 overlaps the detection and
 * event read-back of slice k; a slice holds at least `evals` evaluations over all channels (default 262144; tuning knob). */
syldet_status syldet_batch_set_slice_evals(syldet_batch *b, int64_t evals);

/*
 * Host buffers in, events out (H2D copy, kernels, D2H of the sparse events, host-side debounce).
 *   pcm            planar: channel c starts at pcm + c*channel_stride; interleaved: sample i of channel c at pcm[i*n_channels + c]
 *   pcm_format     SYLDET_PCM_F32 (float), SYLDET_PCM_S16 (int16, converted on the device as x/32768) or SYLDET_PCM_S24 (packed 24-bit, x/2^23)
 *   debounce_frames  TrackDetector.debounceFrames; 0 = every detected evaluation is an event
 *   all_outputs    optional host buffer [n_channels][E][O] receiving every network output (E = num_evals)
 *   events         receives a new syldet_events (sorted by channel, then sample); free with syldet_events_free
 */
syldet_status syldet_batch_run_host(syldet_batch *b, const void *pcm, int pcm_format, int n_channels, int64_t n_samples,
                                    int64_t channel_stride, int layout, int64_t debounce_frames, int detect_rule,
                                    float *all_outputs, syldet_events **events);

/*
 * Simulator trace: what ViewControllerSimulator.simulateNetwork writes next to the input audio
 * (SyllableDetector/ViewControllerSimulator.swift:251-254 initial count, :308-344 per-value fill, :203-211 16-bit LPCM writer).
 * trace is a host buffer [n_channels][n_samples] (planar) of trace_format SYLDET_PCM_F32 or SYLDET_PCM_S16:
 *   trace[s] = 0                                          for s <  first_output_sample
 *            = clamp(out0_j / Float(thresholds[0]), 0, 1)  for s >= first_output_sample, j = (s - first_output_sample) / hop
 *            = 0                                          past the last evaluation's hop (upstream leaves stale buffer contents there)
 * S16 stores min(32767, rint(v * 32768)) and writes NaN (silence through l2normalize) as 0. pcm arguments as for syldet_batch_run_host.
 */
syldet_status syldet_batch_simulate_host(syldet_batch *b, const void *pcm, int pcm_format, int n_channels, int64_t n_samples,
                                         int64_t channel_stride, int layout, int trace_format, void *trace);
/*
 * Device-resident variant: d_pcm is a device pointer (float32), d_all_outputs an optional device buffer.
 * `stream` is a cudaStream_t (NULL = legacy default stream).  Launches are asynchronous; nothing is copied to the host.
 * Call syldet_batch_collect afterwards to synchronise and fetch the events.
 */
syldet_status syldet_batch_launch_device(syldet_batch *b, const float *d_pcm, int n_channels, int64_t n_samples,
                                         int64_t channel_stride, int layout, int detect_rule, float *d_all_outputs,
                                         void *stream);
syldet_status syldet_batch_collect(syldet_batch *b, int64_t debounce_frames, syldet_events **events);
/* Kernels launched by this handle since creation (for bench.py's gpu_launches). */
int64_t syldet_batch_launch_count(const syldet_batch *b);
/* Planning helper (no device needed): tiles per unit the tensor kernel would use for `evals_per_channel` evaluations of each of
 * `n_channels` channels on `sm_count` persistent CTAs with a window of `time_range` columns - the even length that leaves the fewest
 * tiles on the busiest CTA. */
int syldet_plan_tensor_unit_tiles(int64_t evals_per_channel, int n_channels, int sm_count, int time_range);
/* 1 once this handle has switched its tensor kernel to the all-TF32 variant because audio left the fp16 window (see above), else 0. */
int64_t syldet_batch_range_fallbacks(const syldet_batch *b);
/* Device time (ms) of the two kernels of the last SYLDET_KERNEL_WIDE launch, summed over its time segments: the high-overlap STFT
 * (stft_planes_kernel) and the tcgen05 contraction + network tail (wide_l0_kernel). Synchronises. For bench.py's roofline. */
syldet_status syldet_batch_wide_phase_ms(syldet_batch *b, double *stft_ms, double *contraction_ms);
/*
 * CircularShortTimeFourierTransform.extractPower()[f0 ..< f1] (CSTFT.swift:280-337 = |X[k]|, band slice of SyllableDetector.swift:136-148)
 * as the ACTIVE kernel computes it on the detection path, before the spectrogram scaling: band receives [n_channels][n_columns][L]
 * float32 with L = freq_index_range width and *n_columns = num_evals + time_range - 1 (every column that feeds an evaluation; 0 when
 * the recording is shorter than one feature window). Call with band = NULL to query *n_columns. pcm arguments as for
 * syldet_batch_run_host. Inspection / parity-test entry point: the band magnitudes never leave the chip in normal operation.
 * (extractMagnitude(), :221-278, is the square of these values; upstream's detector never calls it.)
 */
syldet_status syldet_batch_spectra_host(syldet_batch *b, const void *pcm, int pcm_format, int n_channels, int64_t n_samples,
                                        int64_t channel_stride, int layout, float *band, int64_t *n_columns);
/* Raw detection count of the last launch before debounce (synchronises). */
syldet_status syldet_batch_last_detection_count(syldet_batch *b, int64_t *count);

int64_t syldet_events_count(const syldet_events *ev);
int syldet_events_outputs_per_event(const syldet_events *ev);
const syldet_event *syldet_events_data(const syldet_events *ev);
const float *syldet_events_outputs(const syldet_events *ev); /* [count][O] */
/* The same rows as separate columns, copied in one pass into caller arrays of syldet_events_count() entries (outputs: [count][O]);
 * any of the three pointers may be NULL. For bindings that want column arrays (numpy, Swift [Int64]). */
void syldet_events_copy_columns(const syldet_events *ev, int32_t *channel, int64_t *sample, float *outputs);
/* The same rows in the packed layout a multi-GPU job gathers (no padding, 12 + 4 O bytes per row): uint32 key = recording << 16 | channel,
 * int64 sample, float outputs[O]; `rows` holds syldet_events_count() of them. Returns 1 when the rows are in (key, sample) order - they
 * are after syldet_batch_run_host / _collect -, 0 otherwise. Recording and channels below 65 536. */
int syldet_events_copy_compact(const syldet_events *ev, uint32_t recording, void *rows);
void syldet_events_free(syldet_events *ev);

/* ---- one stream: class SyllableDetector (Common/SyllableDetector.swift:13-231) ------------------------------------ */
syldet_status syldet_detector_create(const syldet_config *cfg, int device, syldet_detector **out);
void syldet_detector_destroy(syldet_detector *d);
/* appendAudioData(_:withSamples:) (:129-132). SYLDET_ERR_OVERFLOW mirrors the 409 600-byte ring (CSTFT.swift:61,199). */
syldet_status syldet_detector_append(syldet_detector *d, const float *samples, int64_t n);
/* processNewValue() (:153-217): 1 = a new evaluation was produced (see last_outputs), 0 = not enough data, <0 = -status */
int syldet_detector_process_new_value(syldet_detector *d);
syldet_status syldet_detector_last_outputs(const syldet_detector *d, float *out, int cap); /* lastOutputs (:26) */
int syldet_detector_last_detected(const syldet_detector *d);                               /* lastDetected (:27-31) */
int syldet_detector_seen_syllable(syldet_detector *d);                                     /* seenSyllable() (:220-230) */

/* ---- live group: Processor.swift (one detector per channel, small buffers per tick) -------------------------------- */
syldet_status syldet_stream_create(const syldet_config *cfg, int n_channels, int max_buffer, int device, syldet_stream **out);
/*
 * The same group fed at the audio DEVICE's rate: "if abs(config.samplingRate - inputRate) > 1 { resampler = ResamplerLinear(...) }"
 * (SyllableDetector/ViewControllerProcessor.swift:247-250) and the per-channel resampleVector call of receiveAudioFrom
 * (SyllableDetector/Processor.swift:116-121). Every submitted buffer is one ResamplerLinear.resampleVector call per channel
 * (Common/Resampler.swift:35-70, bit-faithful including the carried offset / last sample); the interpolation runs inside the tick
 * kernel, the `last` sample of every channel stays on the device. input_rate within 1 Hz of the configuration's rate: no resampling.
 */
syldet_status syldet_stream_create_resampled(const syldet_config *cfg, int n_channels, int max_buffer, int device, double input_rate,
                                             syldet_stream **out);
int syldet_stream_resampling(const syldet_stream *s); /* 1 when the group resamples its input */
void syldet_stream_destroy(syldet_stream *s);
/*
 * One tick: bufs[c] points at n float32 samples for channel c (receiveAudioFrom(...) for every channel, Processor.swift:102-149).
 * On return seen[c] = 1 iff any evaluation completed by this tick had output 0 >= threshold 0 (the `seen` flag of :131-147),
 * n_new[c] = evaluations completed, last_out[c*O..] = outputs of the newest evaluation (unchanged if none). Arrays may be NULL.
 */
syldet_status syldet_stream_submit(syldet_stream *s, const float *const *bufs, int n, uint8_t *seen, int32_t *n_new,
                                   float *last_out);
int64_t syldet_stream_launch_count(const syldet_stream *s);
/* of those, launches of the latency-shaped tick kernel (configurations the fused plan takes); the rest ran the reference-order tick */
int64_t syldet_stream_fast_tick_count(const syldet_stream *s);
/*
 * Ticks served by the resident tick kernel (SYLDET_STREAM_RESIDENT=1 when the group is created; groups of at most one channel per SM on
 * configurations the latency-shaped tick takes): its blocks stay on their SMs and poll a message in pinned host memory, so such a tick
 * costs no kernel launch. The kernel leaves after SYLDET_STREAM_RESIDENT_IDLE_MS (default 50) without a tick and is started again
 * by the next one.
 */
int64_t syldet_stream_resident_tick_count(const syldet_stream *s);
/*
 * Level meters of the live view: getInputForChannel / getOutputForChannel for every channel (Processor.swift:158-184, StatMax in
 * SummaryStat.swift:39-62). input_rms[c] = sqrt(max over the buffers submitted since the last call of sum(x^2)/n, :110-113),
 * output_max[c] = max over the evaluations since the last call of Double(lastOutputs[0]) (:138); NaN where upstream returns nil.
 * Reads and resets (readStatAndReset). Both maxima are kept on the device by the tick kernel. Either array may be NULL.
 */
syldet_status syldet_stream_read_levels(syldet_stream *s, double *input_rms, double *output_max);
/*
 * TTL-style pulses (ProcessorAudio: Processor.swift:187-221; AudioOutputInterface: AudioInterface.swift:13-40, 442-445).
 * After set_pulse(high_seconds = 0.001 upstream, output_rate), every submit whose `seen` flag is set for a channel arms
 * Int(high_seconds * output_rate) high frames for it; render_pulses fills the next n_frames of every channel's output buffer
 * with 1.0 while armed frames remain and 0.0 after, exactly as the output device's render callback does.
 */
syldet_status syldet_stream_set_pulse(syldet_stream *s, double high_seconds, double output_rate);
syldet_status syldet_stream_render_pulses(syldet_stream *s, float *const *out, int n_frames);

/* ---- ResamplerLinear (Common/Resampler.swift:20-70), bit-faithful including per-buffer state ------------------------ */
syldet_status syldet_resampler_linear_create(double rate_in, double rate_out, syldet_resampler **out);
void syldet_resampler_destroy(syldet_resampler *r);
/* resampleVector(_:ofLength:) (:35-70). *n_out receives the produced count; SYLDET_ERR_ARG if cap is too small. */
syldet_status syldet_resampler_process(syldet_resampler *r, const float *in, int64_t n_in, float *out, int64_t cap,
                                       int64_t *n_out);
/* Upper bound of samples one call can produce for n_in inputs. */
int64_t syldet_resampler_max_output(const syldet_resampler *r, int64_t n_in);

/* ---- batched sample-rate conversion: whole channels at once, on the device ------------------------------------------------
 * LINEAR    = one ResamplerLinear.resampleVector call per channel over the whole buffer, from a fresh state (offset 0, last 0):
 *             n_out = Int(Float(n_in) / Float(rate_in / rate_out)) (Resampler.swift:32,40), float32 index ramp as upstream.
 * POLYPHASE = what upstream gets from AVFoundation for files whose rate differs from the network's (the asset reader is asked for
 *             config.samplingRate, Common/SyllableDetector.swift:19-23): up / down = rate_out / rate_in reduced, low-pass FIR
 *             firwin(20 max(up, down) + 1, 1 / max(up, down), kaiser 5.0) x up, zero-phase, zero padding at both ends,
 *             n_out = ceil(n_in up / down). Integral rates only.
 * in: host, planar, channel c at in + c * in_stride (n_channels == 1: stride ignored). out: host, planar, out_stride >= n_out.
 */
int64_t syldet_resample_output_length(int mode, int64_t n_in, double rate_in, double rate_out); /* -1: unsupported rate pair */
syldet_status syldet_resample_host(int mode, const float *in, int n_channels, int64_t n_in, int64_t in_stride, double rate_in,
                                   double rate_out, float *out, int64_t out_stride, int64_t *n_out, int device);
/* Same on device-resident planar buffers (asynchronous on `stream`, a cudaStream_t; the polyphase filter upload synchronises it once). */
syldet_status syldet_resample_device(int mode, const float *d_in, int n_channels, int64_t n_in, int64_t in_stride, double rate_in,
                                     double rate_out, float *d_out, int64_t out_stride, int64_t *n_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SYLDET_H */
