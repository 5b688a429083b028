// Streaming front-ends (see stream.cu).
#pragma once
#include <chrono>

#include "engine.hpp"
#include "resample.cuh"

namespace syldet {

constexpr int kSampleRingFloats = 409600 / 4;  // CircularShortTimeFourierTransform(buffer: 409600) bytes (CSTFT.swift:61)

class StreamGroup {
public:
    StreamGroup() = default;
    StreamGroup(const StreamGroup &) = delete;
    StreamGroup &operator=(const StreamGroup &) = delete;
    ~StreamGroup();
    // input_rate: sampling rate of the buffers submit() receives. When it differs from the configuration's rate by more than 1 Hz
    // (the rule of ViewControllerProcessor.swift:247-250) every buffer goes through ResamplerLinear inside the tick kernel
    // (Processor.swift:116-121); 0 = the configuration's own rate.
    syldet_status init(const Config &cfg, int n_channels, int max_buffer, int device, double input_rate = 0.0);
    bool resampling() const { return rs_on_; }
    // One tick. *outs -> pinned host [n_channels][*n_new][outputs], valid until the next call.
    syldet_status submit(const float *const *bufs, int n, const float **outs, int64_t *n_new);
    const Config &config() const { return model_.config(); }
    int n_channels() const { return n_channels_; }
    int64_t launch_count() const { return launches_; }
    int64_t fast_tick_count() const { return fast_ticks_; }   // launches of stream_tick_fast_kernel among them
    int64_t resident_tick_count() const { return resident_ticks_; }   // ticks served by the resident kernel (no launch)
    // getInputForChannel / getOutputForChannel for every channel (Processor.swift:158-184): RMS of the loudest buffer and the
    // largest output 0 since the last call, NaN where upstream returns nil; resets both (readStatAndReset).
    syldet_status read_levels(double *input_rms, double *output_max);

private:
    syldet_status wait_for_tick(bool packed);
    syldet_status launch_tick(int64_t n_cols, int64_t avail);
    // resident tick kernel (SYLDET_STREAM_RESIDENT=1): see stream_tick_resident_kernel
    syldet_status start_resident(const StreamTick &t);
    syldet_status stop_resident();
    syldet_status wait_for_resident_tick(const StreamTick &t);

    DeviceModel model_;
    int n_channels_ = 0, max_buffer_ = 0;
    int stage_cap_ = 0;       // floats per channel in the pinned staging area
    int staged_ = 0;          // samples per channel staged on the host, not yet pulled by the device
    int64_t ring_cap_ = 0;    // floats per channel in the device sample ring (power of two)
    int64_t band_cols_ = 0;   // columns per channel in the device band-feature ring (power of two)
    int64_t total_ = 0;       // samples appended per channel so far
    int64_t cols_done_ = 0;   // STFT columns already in the band ring
    int64_t next_eval_ = 0;   // first evaluation not yet produced
    int64_t max_new_ = 0;     // most evaluations one tick can complete
    int64_t launches_ = 0;
    unsigned seq_ = 0;
    DeviceBuffer ring_, band_, counter_, level_in_, level_out_;
    std::vector<int> marks_;  // ends of the buffers waiting in the staging area
    // ResamplerLinear in the tick: phase on the host (data-independent), `last` per channel on the device (two alternating arrays)
    bool rs_on_ = false;
    LinearResamplerPhase rs_;
    int staged_out_ = 0;      // configuration-rate samples the staged buffers will produce
    std::vector<float> mark_offset_;
    std::vector<int> mark_nout_, mark_out0_;
    DeviceBuffer rs_last_;
    int rs_parity_ = 0;
    int64_t buffers_seen_ = 0, evals_seen_ = 0;  // since the last read_levels
    cudaStream_t stream_ = nullptr;
    float *h_stage_ = nullptr, *h_out_ = nullptr;
    unsigned *h_flag_ = nullptr;
    uint4 *h_packed_ = nullptr;  // [n_channels] {out0, out1, out2, seq}: the single-evaluation tick's result, one store per channel
    // SYLDET_STREAM_TIMING=1: device cycle stamps per phase and host microseconds per step, reported by the destructor
    long long *h_stamps_ = nullptr;
    double t_phase_[4] = {0, 0, 0, 0}, t_eval_[6] = {0, 0, 0, 0, 0, 0}, t_host_[3] = {0, 0, 0}, t_sub_[6] = {0, 0, 0, 0, 0, 0};
    int64_t fast_ticks_ = 0;
    bool resident_ok_ = false, resident_running_ = false;
    int resident_blocks_ = 0;         // this group's share of the device's resident blocks (0: launched ticks)
    TickGeom resident_geom_{};
    long long resident_idle_cycles_ = 0;
    void *h_post_ = nullptr;          // pinned: the tick message
    unsigned *h_ctl_ = nullptr;       // pinned: [0] quit (host writes), [16] alive (the dispatcher clears it when the kernel leaves)
    DeviceBuffer mailbox_;            // device: the message as the dispatcher republishes it, then the `leave` word
    int64_t resident_ticks_ = 0, resident_starts_ = 0;
    int64_t t_ticks_ = 0;
    std::chrono::steady_clock::time_point t_submit_{};
};

class Detector {
public:
    syldet_status init(const Config &cfg, int device);
    syldet_status append(const float *samples, int64_t n);
    int process_new_value();  // 1 new value, 0 none, <0 -status
    const std::vector<float> &last_outputs() const { return last_outputs_; }
    bool last_detected() const;
    int seen_syllable();

private:
    StreamGroup group_;
    std::vector<float> pending_;      // appended, not yet sent to the device
    std::vector<float> queue_;        // evaluations computed, not yet handed out
    size_t queue_head_ = 0;
    std::vector<float> last_outputs_;
    int64_t appended_ = 0, cols_extracted_ = 0, evals_returned_ = 0, feature_ring_cols_ = 0;
};

class Resampler {
public:
    Resampler() = default;
    Resampler(const Resampler &) = delete;
    Resampler &operator=(const Resampler &) = delete;
    ~Resampler();
    syldet_status init(double rate_in, double rate_out, int device);
    syldet_status process(const float *in, int64_t n_in, float *out, int64_t cap, int64_t *n_out);
    int64_t max_output(int64_t n_in) const;

private:
    float step_ = 1.0f, last_ = 0.0f, offset_ = 0.0f;  // Resampler.swift:24-26
    int device_ = 0;
    cudaStream_t stream_ = nullptr;
    DeviceBuffer d_in_, d_out_;
};

}  // namespace syldet

struct syldet_stream {
    syldet::StreamGroup g;
    // TTL pulses (ProcessorAudio.prepareOutputFor, Processor.swift:212-221; AudioOutputInterface, AudioInterface.swift:13-40, 442-445)
    int64_t high_frames = 0;         // Int(highDuration * output sample rate); 0 = pulses off
    std::vector<int64_t> high_for;   // outputHighFor[channel]
};
struct syldet_detector {
    syldet::Detector d;
};
struct syldet_resampler {
    syldet::Resampler r;
};
