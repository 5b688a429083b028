// Host-side engine: device copies of a configuration, kernel planning, and the batch runner that stands in for
// TrackDetector.process + the main.swift loop (SyllableDetectorCLI/TrackDetector.swift:45-105, main.swift:126-130).
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "config.hpp"
#include "kernels.hpp"

namespace syldet {

syldet_status cuda_fail(cudaError_t e, const char *what);
#define SYLDET_CUDA(expr)                                                 \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) return ::syldet::cuda_fail(_e, #expr);     \
    } while (0)

// Selects `device` after checking it is a compute-capability 10.x part; there is no CPU fallback.
syldet_status use_device(int device);
int usable_device_count();

// RAII device allocation that grows on demand.
class DeviceBuffer {
public:
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    ~DeviceBuffer();
    syldet_status reserve(size_t bytes);
    void *get() const { return ptr_; }
    template <typename T>
    T *as() const { return static_cast<T *>(ptr_); }
    size_t size() const { return size_; }

private:
    void *ptr_ = nullptr;
    size_t size_ = 0;
};

struct FusedPlan {
    bool ok = false;
    std::string why;  // why the fused kernel does not apply
    FusedParams params{};
    FusedLaunch launch{};
    int blocks_per_sm = 0;
};

// Folds the input-processing chain into layer 0 and packs everything the fused kernel reads from parameter space.
FusedPlan plan_fused(const Config &cfg);

// Tensor-core variant (kernels_tc.cu): same folded parameters, its own ring geometry, plus the windowed DFT matrix.
struct TcPlan {
    bool ok = false;
    std::string why;
    FusedParams params{};
    int hp = 0;
    size_t smem = 0;
    std::vector<float> dft_hi, dft_lo;  // [128][k_pad]
    std::vector<uint32_t> dft16;        // [128][tc_a16_cols()] fp16 pairs: two-term fp16 split of the DFT matrix (kF16 variant, kernels_tc.cu)
    std::vector<float> wcat_hi, wcat_lo;  // [n0][32]
    int n0 = 0;
};
TcPlan plan_tc(const Config &cfg, const FusedPlan &fused);

// Wide-hidden tensor path (kernels_wide.cu): two-layer networks on a hop-4 STFT; layer 0 as a 3xTF32 tcgen05 contraction.
struct WidePlan {
    bool ok = false;
    std::string why;
    WideParams params{};
    std::vector<float> weights;   // blocks [pass][chunk][t][hi | lo][plane][256][4] of the folded layer-0 weights (tf32 hi / lo)
    std::vector<float> v, bprime; // [h_pad]
    std::vector<float> w1;        // [n_out][h_pad]
};
WidePlan plan_wide(const Config &cfg);

// Device-resident copy of one configuration (weights, window, twiddles, DevNet record).
class DeviceModel {
public:
    DeviceModel() = default;
    DeviceModel(const DeviceModel &) = delete;
    DeviceModel &operator=(const DeviceModel &) = delete;
    syldet_status init(const Config &cfg, int device);
    const Config &config() const { return cfg_; }
    int device() const { return device_; }
    const DevNet *dev_net() const { return d_net_.as<DevNet>(); }
    const float *window() const { return d_window_; }
    const float2 *twiddle() const { return d_twiddle_; }
    int max_width() const { return max_width_; }
    const FusedPlan &fused() const { return fused_; }
    const TcPlan &tc() const { return tc_; }
    const WidePlan &wide() const { return wide_; }
    const float *wide_weights() const { return d_wide_.as<float>(); }
    const float *wide_v() const { return d_wide_.as<float>() + wide_.weights.size(); }
    const float *wide_bprime() const { return wide_v() + wide_.v.size(); }
    const float *wide_w1() const { return wide_bprime() + wide_.bprime.size(); }
    const float *dft_hi() const { return d_dft_.as<float>(); }
    const float *dft_lo() const { return d_dft_.as<float>() + 128 * (size_t)tc_k_pad(); }
    const float *wcat_hi() const { return d_dft_.as<float>() + 2 * 128 * (size_t)tc_k_pad(); }
    const float *wcat_lo() const { return wcat_hi() + (size_t)tc_.n0 * 32; }
    const uint32_t *dft16() const { return reinterpret_cast<const uint32_t *>(wcat_lo() + (size_t)tc_.n0 * 32); }
    int sm_count() const { return sm_count_; }
    const FusedParams *fused_params_dev() const { return d_fused_params_.as<FusedParams>(); }   // device copy (live tick: lane-parallel weight reads)
    const unsigned char *blob() const { return d_blob_.as<unsigned char>(); }
    size_t blob_bytes() const { return blob_bytes_; }

private:
    Config cfg_;
    int device_ = -1, max_width_ = 0, sm_count_ = 148;
    DeviceBuffer d_blob_, d_net_, d_dft_, d_wide_, d_fused_params_;
    WidePlan wide_;
    size_t blob_bytes_ = 0;
    TcPlan tc_;
    const float *d_window_ = nullptr;
    const float2 *d_twiddle_ = nullptr;
    FusedPlan fused_;
};

// Header of the device-side event sink: detection count (may exceed the capacity) and the range flag the tensor kernel raises when
// its fp16 correction pass meets audio outside the window in which it is at float32 level (kernels_tc.cu).
struct SinkHeader {
    unsigned long long count;
    int range_flag;
    int reserved;
};
constexpr size_t kSinkHeaderBytes = sizeof(SinkHeader);
// Lower bound of the window's band energy sum |X|^2 (= squared norm of the network input before l2normalize) for the fp16 correction
// pass: its absolute operand errors (2^-25 per subnormal sample, x |A_lo| <= 2^-12) add ~6e-11 per band magnitude, i.e. less than 1e-6
// to a network output while the norm is >= 2^-8 (audio rms >~ 2e-5). Quieter float audio takes the all-TF32 variant.
constexpr float kTcGuardLo = 1.52587890625e-05f;   // 2^-16
constexpr float kTcGuardRange = 0.00390625f;       // 2^-8: the same bound for a window's max - min (min / max normalised windows)

struct EventKey {   // channel << 40 | evaluation, and where the event sits in the sink
    uint64_t key;
    uint32_t idx;
};

struct Events {
    int outputs_per_event = 0;
    std::vector<syldet_event> rows;
    std::vector<float> outputs;
};

class Batch {
public:
    syldet_status init(const Config &cfg, int device);
    ~Batch();
    syldet_status set_kernel(int kernel);
    int active_kernel() const;
    syldet_status run_host(const void *pcm, int fmt, int n_channels, int64_t n_samples, int64_t ch_stride, int layout,
                           int64_t debounce_frames, int detect_rule, float *all_outputs, Events &out, int trace_format = 0,
                           void *trace = nullptr);
    syldet_status launch_device(const float *d_pcm, int n_channels, int64_t n_samples, int64_t ch_stride, int layout,
                                int detect_rule, float *d_all_outputs, cudaStream_t stream, bool pcm_exact = false);
    syldet_status collect(int64_t debounce_frames, Events &out);
    syldet_status last_detection_count(int64_t *count);
    // launches repeated with the all-TF32 variant because the fp16 range flag went up (at most 1: the switch is permanent)
    int64_t range_fallbacks() const { return range_fallbacks_; }
    // device time of the last wide-path launch (all time segments): STFT-planes kernel, contraction + epilogue kernel (synchronises)
    syldet_status wide_phase_ms(double *stft_ms, double *contraction_ms);
    // band magnitudes extractPower()[f0 ..< f1] of every column that feeds an evaluation, from the active kernel: [n_channels][E + T - 1][band]
    syldet_status spectra_host(const void *pcm, int fmt, int n_channels, int64_t n_samples, int64_t ch_stride, int layout, float *band,
                               int64_t *n_columns);
    int64_t launch_count() const { return launches_; }
    void set_slice_evals(int64_t evals) { slice_evals_ = evals; }
    void set_debug_band(float *d_band, int64_t cols) { debug_band_ = d_band; debug_cols_ = cols; }
    const DeviceModel &model() const { return model_; }

private:
    syldet_status launch_planar(const float *d_planar, int n_channels, int64_t n_samples, int64_t ch_stride,
                                const float *valid_begin, const float *valid_end, int detect_rule, float *d_all_outputs,
                                cudaStream_t stream);
    syldet_status ensure_sink(unsigned long long capacity);
    syldet_status launch_fused_range(const float *d_planar, int n_channels, int64_t ch_stride, const float *valid_begin,
                                     const float *valid_end, int64_t eval_begin, int64_t eval_count, int64_t evals_total,
                                     int detect_rule, float *d_all_outputs, EventSink sink, cudaStream_t stream);
    syldet_status launch_tc_range(const float *d_planar, int n_channels, int64_t n_samples, int64_t ch_stride, int64_t eval_offset,
                                  int64_t eval_count, int64_t evals_total, int detect_rule, float *d_all_outputs, EventSink sink,
                                  cudaStream_t stream, const int16_t *d_s16 = nullptr);   // d_s16: the same position in planar 16-bit PCM (ch_stride then counts its samples)
    syldet_status launch_planar_range(const float *d_planar, int n_channels, int64_t n_samples, int64_t n_avail, int64_t ch_stride,
                                      const float *valid_begin, const float *valid_end, int64_t eval_begin, int64_t eval_count,
                                      int detect_rule, float *d_all_outputs, bool reset_sink, cudaStream_t stream);
    syldet_status ensure_pipeline(int slices, size_t event_bytes);
    syldet_status settle(unsigned long long *n_events);
    syldet_status order_events_on_device(unsigned long long n, int64_t evals, int n_channels, cudaStream_t stream, bool *ordered);
    syldet_status launch_wide_range(const float *d_planar, int n_channels, int64_t ch_stride, int64_t eval_begin, int64_t eval_count,
                                    int64_t evals_total, int detect_rule, float *d_all_outputs, EventSink sink, cudaStream_t stream);

    DeviceModel model_;
    int kernel_ = SYLDET_KERNEL_AUTO;
    cudaStream_t own_stream_ = nullptr;
    // run_host pipeline: time slices are copied on copy_stream_, detected on own_stream_, their events read back on d2h_stream_
    cudaStream_t copy_stream_ = nullptr, d2h_stream_ = nullptr;
    std::vector<cudaEvent_t> ev_copied_, ev_done_;
    SinkHeader *h_counts_ = nullptr;           // pinned, one snapshot of the sink header (running event count, range flag) per slice
    bool f16_ok_ = true;                       // false once the range flag went up: the tensor kernel runs all-TF32 from then on
    bool pcm_exact_ = false;                   // the current input came from 16-bit PCM (exact fp16 operands)
    int64_t range_fallbacks_ = 0;
    void *h_events_ = nullptr;                 // pinned: DevEvent[capacity] then float[capacity][outputs]
    size_t h_events_bytes_ = 0;
    int64_t slice_evals_ = 256 * 1024;
    DeviceBuffer planar_, feat_, sink_count_, sink_events_, sink_outputs_, staging_;
    DeviceBuffer order_bits_, order_prefix_, order_blocks_, sorted_events_, sorted_outputs_;   // device-side ordering of the detections (collect)
    std::vector<EventKey> collect_keys_;   // scratch of collect(), kept between calls
    size_t max_pitch_ = 0;                 // cudaDevAttrMaxPitch: largest pitch cudaMemcpy2D accepts
    size_t last_event_count_ = 0;          // events of the previous run_host (+ margin): how much of the result to touch up front
    std::vector<cudaEvent_t> wide_ev_;
    int wide_segments_ = 0;
    DeviceBuffer wide_hi_, wide_lo_, wide_stats_;   // band-magnitude planes + column statistics of one time segment (wide path)
    unsigned long long sink_capacity_ = 0;
    int64_t launches_ = 0;
    float *debug_band_ = nullptr;
    int64_t debug_cols_ = 0;
    // last launch, kept so that an event-buffer overflow can be replayed with a larger buffer
    struct Last {
        bool valid = false;
        const float *d_pcm = nullptr;
        int n_channels = 0;
        int64_t n_samples = 0, ch_stride = 0;
        int layout = 0, detect_rule = 0;
        float *d_all_outputs = nullptr;
        cudaStream_t stream = nullptr;
        bool pcm_exact = false;
    } last_;
};

// Batched sample-rate conversion (resample.cu).
int64_t resample_output_length(int mode, int64_t n_in, double rate_in, double rate_out);
syldet_status resample_device(int mode, const float *d_in, int n_channels, int64_t n_in, int64_t in_stride, double rate_in, double rate_out,
                              float *d_out, int64_t out_stride, int64_t *n_out, DeviceBuffer &filter, cudaStream_t stream);
syldet_status resample_host(int mode, const float *in, int n_channels, int64_t n_in, int64_t in_stride, double rate_in, double rate_out, float *out,
                            int64_t out_stride, int64_t *n_out, int device);

// Greedy debounce over ascending sample numbers of one channel (TrackDetector.swift:80,99).
void debounce_sorted(const Config &cfg, std::vector<syldet_event> &rows, std::vector<float> &outputs, int n_out,
                     int64_t debounce_frames);

}  // namespace syldet

struct syldet_batch {
    syldet::Batch b;
};
struct syldet_events {
    syldet::Events e;
};
