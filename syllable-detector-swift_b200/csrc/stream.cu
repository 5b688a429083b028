// Streaming front-ends on top of the batch kernels:
//   StreamGroup  - Processor.swift's shape: many live channels, one small buffer per channel per tick
//                  (SyllableDetector/Processor.swift:102-149, AudioInterface.swift:474 frameSize 32)
//   Detector     - class SyllableDetector, one stream with appendAudioData / processNewValue / lastOutputs
//                  (Common/SyllableDetector.swift:13-231)
//   Resampler    - ResamplerLinear (Common/Resampler.swift:20-70), the index ramp and interpolation run in a kernel
// Per channel the device keeps a circular sample history and a circular band-feature history, so a tick computes only
// the STFT columns and evaluations the new samples complete. A tick is ONE kernel launch (stream_tick_kernel): the
// kernel pulls the new samples from pinned host memory itself, writes the outputs back into pinned host memory and
// raises a host-visible flag - no memcpy calls and no stream synchronisation on the latency path. Ticks that complete
// no STFT column (3 of 4 at 32-frame buffers and hop 132) only stage samples on the host.
#include "stream.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>

namespace syldet {

// ---------------------------------------------------------------------------------------------------------------
namespace {
// Resident tick kernels of this process, in blocks per device: their blocks wait for each other through memory, so all of them - of
// every group - have to fit on the device's SMs at once. A group that would not fit keeps to launched ticks.
std::atomic<int> g_resident_blocks[64];

int64_t pow2_at_least(int64_t v) {
    int64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
}  // namespace

syldet_status StreamGroup::init(const Config &cfg, int n_channels, int max_buffer, int device, double input_rate) {
    if (n_channels <= 0 || n_channels > 65535) return set_error(SYLDET_ERR_ARG, "n_channels out of range");
    if (max_buffer <= 0) return set_error(SYLDET_ERR_ARG, "max_buffer must be positive");
    if (input_rate < 0.0 || input_rate != input_rate) return set_error(SYLDET_ERR_ARG, "input rate must be positive (or 0)");
    syldet_status st = model_.init(cfg, device);
    if (st != SYLDET_OK) return st;
    n_channels_ = n_channels;
    max_buffer_ = max_buffer;
    const Config &c = model_.config();
    if (stream_tick_smem(c.fourier_length, model_.max_width(), nullptr) > kStreamTickMaxSmem)
        return set_error(SYLDET_ERR_CONFIG, "configuration too large for the live tick kernel");
    // Samples wait in pinned staging until they complete an STFT column: fewer than max(W + gap, hop) can be waiting
    // when a buffer arrives (see submit), so this never overflows.
    const int frame = c.gap + c.window_length;
    // "if abs(config.samplingRate - inputRate) > 1 { resampler = ResamplerLinear(fromRate:toRate:) }" (ViewControllerProcessor.swift:247-250)
    rs_on_ = input_rate > 0.0 && std::fabs(c.sampling_rate - input_rate) > 1.0;
    double in_per_out = 1.0;
    if (rs_on_) {
        rs_.step = (float)(input_rate / c.sampling_rate);   // Resampler.swift:32
        rs_.offset = 0.0f;
        in_per_out = (double)rs_.step;
    }
    // staging holds device-rate samples: what can wait before a column completes, converted to the device rate (+ one sample per
    // buffer for the resampler's carry), plus one more buffer; the ring holds configuration-rate samples
    const int wait_out = std::max(frame, c.hop);
    const int64_t wait_in = rs_on_ ? (int64_t)std::ceil(wait_out * in_per_out) + 2 * kStreamMaxMarks + 8 : wait_out;
    stage_cap_ = (int)(((wait_in + max_buffer + 31) / 32) * 32);
    const int64_t out_cap = rs_on_ ? (int64_t)std::ceil(stage_cap_ / in_per_out) + kStreamMaxMarks + 8 : stage_cap_;
    ring_cap_ = pow2_at_least((int64_t)frame + out_cap);
    max_new_ = out_cap / c.hop + 2;
    band_cols_ = pow2_at_least(c.time_range + max_new_);
    st = ring_.reserve((size_t)n_channels * ring_cap_ * sizeof(float));
    if (st != SYLDET_OK) return st;
    st = band_.reserve((size_t)n_channels * band_cols_ * c.band * sizeof(float));
    if (st != SYLDET_OK) return st;
    st = counter_.reserve(sizeof(unsigned));
    if (st != SYLDET_OK) return st;
    st = level_in_.reserve((size_t)n_channels * sizeof(unsigned long long));
    if (st != SYLDET_OK) return st;
    st = level_out_.reserve((size_t)n_channels * sizeof(int));
    if (st != SYLDET_OK) return st;
    st = rs_last_.reserve((size_t)2 * n_channels * sizeof(float));
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    SYLDET_CUDA(cudaMemsetAsync(rs_last_.get(), 0, rs_last_.size(), stream_));   // var last: Float = 0.0 (Resampler.swift:25)
    SYLDET_CUDA(cudaMemsetAsync(ring_.get(), 0, ring_.size(), stream_));
    SYLDET_CUDA(cudaMemsetAsync(band_.get(), 0, band_.size(), stream_));
    SYLDET_CUDA(cudaMemsetAsync(counter_.get(), 0, sizeof(unsigned), stream_));
    SYLDET_CUDA(cudaMemsetAsync(level_in_.get(), 0, level_in_.size(), stream_));
    SYLDET_CUDA(cudaMemsetAsync(level_out_.get(), 0x80, level_out_.size(), stream_));  // 0x80808080: below the image of every finite float
    // cudaMallocHost memory is device-visible at the same address under unified addressing (always on for sm_100)
    SYLDET_CUDA(cudaMallocHost(&h_stage_, (size_t)n_channels * stage_cap_ * sizeof(float)));
    SYLDET_CUDA(cudaMallocHost(&h_out_, (size_t)n_channels * max_new_ * c.outputs * sizeof(float)));
    SYLDET_CUDA(cudaMallocHost(&h_flag_, (size_t)n_channels * sizeof(unsigned)));
    std::memset(h_flag_, 0, (size_t)n_channels * sizeof(unsigned));
    SYLDET_CUDA(cudaMallocHost(&h_packed_, (size_t)n_channels * sizeof(uint4)));
    std::memset(h_packed_, 0, (size_t)n_channels * sizeof(uint4));
    if (const char *e = std::getenv("SYLDET_STREAM_TIMING"); e && e[0] == '1') {
        SYLDET_CUDA(cudaMallocHost(&h_stamps_, 128));
        std::memset(h_stamps_, 0, 128);
    }
    SYLDET_CUDA(cudaStreamSynchronize(stream_));
    // Resident tick kernel (opt-in, SYLDET_STREAM_RESIDENT=1): one polling block per channel, all co-resident
    if (const char *e = std::getenv("SYLDET_STREAM_RESIDENT"); e && e[0] == '1') {
        const FusedPlan &fp = model_.fused();
        std::atomic<int> &in_use = g_resident_blocks[device & 63];
        const bool room = in_use.fetch_add(n_channels + 1) + n_channels + 1 <= model_.sm_count();
        if (!room) in_use.fetch_sub(n_channels + 1);
        if (room && !(fp.ok && stream_tick_fast_supported(c.fourier_length, fp.params) &&
                      stream_tick_resident_plan(c.fourier_length, fp.launch.hp, fp.params, stage_cap_, &resident_geom_))) {
            in_use.fetch_sub(n_channels + 1);
        } else if (room) {
            resident_blocks_ = n_channels + 1;
            SYLDET_CUDA(cudaMallocHost(&h_post_, stream_tick_post_bytes()));
            std::memset(h_post_, 0, stream_tick_post_bytes());
            SYLDET_CUDA(cudaMallocHost(&h_ctl_, 32 * sizeof(unsigned)));
            std::memset(h_ctl_, 0, 32 * sizeof(unsigned));
            st = mailbox_.reserve(stream_tick_post_bytes() + 64);
            if (st != SYLDET_OK) return st;
            double idle_ms = 50.0;   // a group nobody feeds leaves the GPU after this long; the next tick starts the kernel again
            if (const char *m = std::getenv("SYLDET_STREAM_RESIDENT_IDLE_MS")) idle_ms = std::max(0.1, std::atof(m));
            int khz = 0;
            cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
            resident_idle_cycles_ = (long long)(idle_ms * (khz > 0 ? khz : 1900000));
            resident_ok_ = true;
        }
    }
    return SYLDET_OK;
}

syldet_status StreamGroup::start_resident(const StreamTick &t) {
    const Config &c = model_.config();
    const FusedPlan &fp = model_.fused();
    SYLDET_CUDA(cudaMemsetAsync(mailbox_.get(), 0, mailbox_.size(), stream_));
    *(volatile unsigned *)h_ctl_ = 0;
    *(volatile unsigned *)(h_ctl_ + 16) = 1;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    unsigned char *mb = mailbox_.as<unsigned char>();
    StreamTick first = t;
    first.packed = h_packed_;   // which of `packed` / `out` + `flags` a tick publishes through is part of its message
    // `t` is the tick about to be posted: the kernel starts at its sequence number and keeps its constant part
    SYLDET_CUDA(launch_stream_tick_resident(c.fourier_length, fp.launch.hp, model_.fused_params_dev(), h_post_, h_ctl_, h_ctl_ + 16, mb,
                                            reinterpret_cast<unsigned *>(mb + stream_tick_post_bytes()), first, resident_geom_,
                                            resident_idle_cycles_, model_.window(), model_.twiddle(), n_channels_, model_.sm_count(), stream_));
    resident_running_ = true;
    ++resident_starts_;
    ++launches_;
    return SYLDET_OK;
}

syldet_status StreamGroup::stop_resident() {
    if (!resident_running_) return SYLDET_OK;
    *(volatile unsigned *)h_ctl_ = 1;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    resident_running_ = false;
    SYLDET_CUDA(cudaStreamSynchronize(stream_));
    return SYLDET_OK;
}

// As wait_for_tick, but the kernel may have left (idle limit reached just as the message was posted): the dispatcher takes a tick for
// every channel or for none, so the kernel is simply started again at the pending tick.
syldet_status StreamGroup::wait_for_resident_tick(const StreamTick &t) {
    const bool packed = t.packed != nullptr;
    volatile unsigned *flags = packed ? &h_packed_[0].w : h_flag_;
    const int step = packed ? 4 : 1;
    int ch = 0;
    for (unsigned spins = 0;; ++spins) {
        while (ch < n_channels_ && flags[(size_t)ch * step] == seq_) ++ch;
        if (ch == n_channels_) return SYLDET_OK;
        if ((spins & 0xfff) == 0xfff) {
            if (*(volatile unsigned *)(h_ctl_ + 16) == 0) {   // it left without taking this tick: start it again at the pending one
                syldet_status st = stop_resident();
                if (st != SYLDET_OK) return st;
                st = start_resident(t);
                if (st != SYLDET_OK) return st;
            } else if ((spins & 0xffff) == 0xffff) {
                cudaError_t e = cudaStreamQuery(stream_);
                if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "resident live tick");
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_submit_).count() > 5.0) {
                    // e.g. another process holds the SMs its blocks need: fail loudly rather than spin
                    stop_resident();
                    return set_error(SYLDET_ERR_CUDA, "the resident tick kernel did not answer within 5 s");
                }
            }
        }
    }
}

StreamGroup::~StreamGroup() {
    if (resident_running_) {
        cudaSetDevice(model_.device());
        stop_resident();
    }
    if (resident_blocks_ > 0) g_resident_blocks[model_.device() & 63].fetch_sub(resident_blocks_);
    if (h_post_) cudaFreeHost(h_post_);
    if (h_ctl_) cudaFreeHost(h_ctl_);
    if (h_stamps_ && t_ticks_ > 0) {
        const double n = (double)t_ticks_;
        std::fprintf(stderr, "[syldet stream timing] %lld single-launch ticks, %d channels; device cycles of block (0,0): preload %.0f, copy %.0f, "
                             "columns %.0f, evaluations %.0f (gather %.0f, input processing %.0f + %.0f, layers %.0f + %.0f, output %.0f); host us: stage %.2f, launch call %.2f, wait %.2f\n",
                     (long long)t_ticks_, n_channels_, t_phase_[0] / n, t_phase_[1] / n, t_phase_[2] / n, t_phase_[3] / n,
                     t_eval_[0] / n, t_eval_[1] / n, t_eval_[2] / n, t_eval_[3] / n, t_eval_[4] / n, t_eval_[5] / n,
                     t_host_[0] / n, t_host_[1] / n, t_host_[2] / n);
        if (t_sub_[5] > 0)
            std::fprintf(stderr, "[syldet stream timing] latency-shaped tick, cycles since entry: meters %.0f, pass 1 %.0f, pass 2 %.0f, untangle %.0f, layer 0 %.0f, tail %.0f\n",
                         t_sub_[0] / n, t_sub_[1] / n, t_sub_[2] / n, t_sub_[3] / n, t_sub_[4] / n, t_sub_[5] / n);
    }
    if (h_stamps_) cudaFreeHost(h_stamps_);
    if (stream_) {
        cudaSetDevice(model_.device());
        cudaStreamSynchronize(stream_);
        cudaStreamDestroy(stream_);
    }
    if (h_stage_) cudaFreeHost(h_stage_);
    if (h_out_) cudaFreeHost(h_out_);
    if (h_flag_) cudaFreeHost(h_flag_);
    if (h_packed_) cudaFreeHost(h_packed_);
}

// Every channel's block stores the tick's sequence number into pinned host memory once its outputs are visible (for a
// single evaluation of <= 3 outputs: in the same 16-byte store as the outputs, so no system-scope fence is needed); polling those words is cheaper than cudaStreamSynchronize and needs no device-side counting. The stream is queried now and then so a failed launch cannot spin forever.
syldet_status StreamGroup::wait_for_tick(bool packed) {
    volatile unsigned *flags = packed ? &h_packed_[0].w : h_flag_;
    const int step = packed ? 4 : 1;  // words between consecutive channels
    int ch = 0;  // channels [0, ch) have published this tick
    for (unsigned spins = 0;; ++spins) {
        while (ch < n_channels_ && flags[(size_t)ch * step] == seq_) ++ch;
        if (ch == n_channels_) return SYLDET_OK;
        if ((spins & 0x3fff) == 0x3fff) {
            cudaError_t e = cudaStreamQuery(stream_);
            if (e == cudaSuccess) {
                while (ch < n_channels_ && flags[(size_t)ch * step] == seq_) ++ch;
                if (ch == n_channels_) return SYLDET_OK;
                return set_error(SYLDET_ERR_CUDA, "live tick finished without publishing its results");
            }
            if (e != cudaErrorNotReady) return cuda_fail(e, "live tick");
        }
    }
}

syldet_status StreamGroup::submit(const float *const *bufs, int n, const float **outs, int64_t *n_new) {
    t_submit_ = std::chrono::steady_clock::now();
    *n_new = 0;
    *outs = h_out_;
    if (n < 0 || n > max_buffer_) return set_error(SYLDET_ERR_ARG, "buffer longer than max_buffer");
    if (n == 0) return SYLDET_OK;
    const Config &c = model_.config();
    if (staged_ + n > stage_cap_) return set_error(SYLDET_ERR_OVERFLOW, "Insufficient space on buffer.");
    for (int ch = 0; ch < n_channels_; ++ch)
        std::memcpy(h_stage_ + (size_t)ch * stage_cap_ + staged_, bufs[ch], (size_t)n * sizeof(float));
    int64_t n_out = n;
    if (rs_on_) {   // one resampleVector call per buffer (Processor.swift:116-121); the phase is data-independent, so it is kept here
        n_out = std::max<int64_t>(0, rs_.plan(n));
        mark_offset_.push_back(rs_.offset);
        mark_nout_.push_back((int)n_out);
        mark_out0_.push_back(staged_out_);
        rs_.advance(n, n_out);
    }
    staged_ += n;
    staged_out_ += (int)n_out;
    total_ += n_out;
    marks_.push_back(staged_);
    ++buffers_seen_;
    const int64_t n_cols = c.num_columns(total_) - cols_done_;
    // no column completed: the decision (nothing new) needs no device work; the samples wait in the staging area
    if (n_cols <= 0 && (int)marks_.size() < kStreamMaxMarks) return SYLDET_OK;
    const int64_t avail = c.num_evals(total_) - next_eval_;
    if (avail > max_new_ || n_cols > band_cols_ - c.time_range)
        return set_error(SYLDET_ERR_OVERFLOW, "more evaluations pending than the stream was sized for");
    syldet_status st = launch_tick(n_cols, avail);
    if (st != SYLDET_OK) return st;
    *n_new = avail;
    return SYLDET_OK;
}

// Sends everything in the staging area through the device: copy into the sample rings (+ level meter), the n_cols STFT columns
// and the `avail` evaluations it completes; returns when the results are host-visible.
syldet_status StreamGroup::launch_tick(int64_t n_cols, int64_t avail) {
    const Config &c = model_.config();
    syldet_status st = use_device(model_.device());
    if (st != SYLDET_OK) return st;
    StreamTick t{};
    t.staged = h_stage_;
    t.stage_pitch = stage_cap_;
    t.n_staged = staged_;
    t.ring = ring_.as<float>();
    t.ring_mask = ring_cap_ - 1;
    t.ring_pos = total_ - staged_out_;
    t.rs_on = rs_on_ ? 1 : 0;
    if (rs_on_) {
        t.rs_step = rs_.step;
        t.rs_last_in = rs_last_.as<float>() + (size_t)rs_parity_ * n_channels_;
        t.rs_last_out = rs_last_.as<float>() + (size_t)(rs_parity_ ^ 1) * n_channels_;
        for (size_t k = 0; k < marks_.size(); ++k) {
            t.rs_offset[k] = mark_offset_[k];
            t.rs_n_out[k] = mark_nout_[k];
            t.rs_out0[k] = mark_out0_[k];
        }
    }
    t.band = band_.as<float>();
    t.band_mask = band_cols_ - 1;
    t.col0 = cols_done_;
    t.n_cols = n_cols;
    t.eval0 = next_eval_;
    t.n_evals = avail;
    t.out = h_out_;
    t.counter = counter_.as<unsigned>();
    t.seq = ++seq_;
    t.stamps = h_stamps_;
    t.blob = model_.blob();
    t.blob_bytes = (int)model_.blob_bytes();
    t.level_in = level_in_.as<unsigned long long>();
    t.level_out = level_out_.as<int>();
    t.n_marks = (int)marks_.size();
    for (int k = 0; k < t.n_marks; ++k) t.marks[k] = marks_[k];
    int warps = 4;
    stream_tick_smem(c.fourier_length, model_.max_width(), &warps);
    const DevNet *net = model_.dev_net();
    const int64_t cap = std::max<int64_t>(1, (model_.sm_count() * 8) / n_channels_);
    const bool single = n_cols <= warps;
    bool resident = false;
    const auto tp1 = std::chrono::steady_clock::now();
    if (single) {  // the live shape: one launch, one block per channel
        t.phases = STREAM_PHASE_COPY | STREAM_PHASE_COLUMNS | (avail > 0 ? STREAM_PHASE_EVALS : 0);
        t.flags = h_flag_;
        if (avail == 1 && c.outputs <= 3) t.packed = h_packed_;
        // configurations the fused kernel takes: register FFT + folded network (stream_tick_fast_kernel); SYLDET_STREAM_GENERIC=1 keeps
        // the reference-order tick
        static const bool generic_tick = std::getenv("SYLDET_STREAM_GENERIC") != nullptr;
        const FusedPlan &fp = model_.fused();
        if (resident_ok_ && stream_tick_resident_tick_fits(resident_geom_, fp.params, t)) {
            // no launch: post the tick to the blocks already on the SMs
            if (resident_running_ && *(volatile unsigned *)(h_ctl_ + 16) == 0) {   // idle limit reached since the last tick
                st = stop_resident();
                if (st != SYLDET_OK) return st;
            }
            if (!resident_running_) {
                st = start_resident(t);
                if (st != SYLDET_OK) return st;
            }
            stream_tick_post_write(h_post_, t);
            resident = true;
            ++resident_ticks_;
        } else if (!generic_tick && fp.ok && stream_tick_fast_supported(c.fourier_length, fp.params) &&
            stream_tick_fast_fits(c.fourier_length, fp.launch.hp, fp.params, t)) {
            st = stop_resident();
            if (st != SYLDET_OK) return st;
            SYLDET_CUDA(launch_stream_tick_fast(c.fourier_length, fp.launch.hp, fp.params, model_.fused_params_dev(), t, model_.window(),
                                                model_.twiddle(), n_channels_, stream_));
            ++fast_ticks_;
        } else {
            st = stop_resident();
            if (st != SYLDET_OK) return st;
            SYLDET_CUDA(launch_stream_tick(net, c.fourier_length, model_.max_width(), n_channels_, 1, t, stream_));
        }
        if (!resident) ++launches_;
    } else {  // a long buffer: one launch per phase so each can spread over many blocks per channel
        st = stop_resident();
        if (st != SYLDET_OK) return st;
        t.phases = STREAM_PHASE_COPY;
        t.flags = nullptr;
        int bx = (int)std::min<int64_t>(cap, (staged_ + warps * 32 - 1) / (warps * 32));
        SYLDET_CUDA(launch_stream_tick(net, c.fourier_length, model_.max_width(), n_channels_, bx, t, stream_));
        t.phases = STREAM_PHASE_COLUMNS;
        t.flags = avail > 0 ? nullptr : h_flag_;
        bx = (int)std::min<int64_t>(cap, (n_cols + warps - 1) / warps);
        SYLDET_CUDA(launch_stream_tick(net, c.fourier_length, model_.max_width(), n_channels_, bx, t, stream_));
        launches_ += 2;
        if (avail > 0) {
            t.phases = STREAM_PHASE_EVALS;
            t.flags = h_flag_;
            bx = (int)std::min<int64_t>(cap, (avail + warps - 1) / warps);
            SYLDET_CUDA(launch_stream_tick(net, c.fourier_length, model_.max_width(), n_channels_, bx, t, stream_));
            ++launches_;
        }
    }
    const auto tp2 = std::chrono::steady_clock::now();
    st = resident ? wait_for_resident_tick(t) : wait_for_tick(t.packed != nullptr);
    if (st != SYLDET_OK) return st;
    if (t.packed) {  // unpack into the documented [n_channels][n_new][outputs] layout
        for (int ch = 0; ch < n_channels_; ++ch) std::memcpy(h_out_ + (size_t)ch * c.outputs, &h_packed_[ch], (size_t)c.outputs * sizeof(float));
    }
    if (h_stamps_ && single && avail > 0) {
        const auto tp3 = std::chrono::steady_clock::now();
        for (int k = 0; k < 4; ++k) t_phase_[k] += (double)(h_stamps_[k + 1] - h_stamps_[k]);
        t_eval_[0] += (double)(h_stamps_[5] - h_stamps_[3]);   // gather
        for (int k = 1; k < 5; ++k) t_eval_[k] += (double)(h_stamps_[5 + k] - h_stamps_[4 + k]);  // ip0, ip1, layer0, layer1
        t_eval_[5] += (double)(h_stamps_[4] - h_stamps_[9]);   // reverse maps + publish
        for (int k = 0; k < 6; ++k) t_sub_[k] += (double)(h_stamps_[10 + k] - h_stamps_[0]);   // latency-shaped tick: cycles since entry
        t_host_[0] += std::chrono::duration<double, std::micro>(tp1 - t_submit_).count();
        t_host_[1] += std::chrono::duration<double, std::micro>(tp2 - tp1).count();
        t_host_[2] += std::chrono::duration<double, std::micro>(tp3 - tp2).count();
        ++t_ticks_;
    }
    staged_ = 0;
    staged_out_ = 0;
    marks_.clear();
    if (rs_on_) {
        mark_offset_.clear();
        mark_nout_.clear();
        mark_out0_.clear();
        rs_parity_ ^= 1;
    }
    cols_done_ += n_cols;
    next_eval_ += avail;
    evals_seen_ += avail;
    return SYLDET_OK;
}

syldet_status StreamGroup::read_levels(double *input_rms, double *output_max) {
    if (staged_ > 0) {  // buffers still waiting on the host count too
        syldet_status st = launch_tick(0, 0);
        if (st != SYLDET_OK) return st;
    }
    syldet_status st = use_device(model_.device());
    if (st != SYLDET_OK) return st;
    st = stop_resident();   // the copies below queue behind whatever runs on the stream
    if (st != SYLDET_OK) return st;
    std::vector<unsigned long long> in(n_channels_);
    std::vector<int> out(n_channels_);
    SYLDET_CUDA(cudaMemcpyAsync(in.data(), level_in_.get(), in.size() * sizeof(in[0]), cudaMemcpyDeviceToHost, stream_));
    SYLDET_CUDA(cudaMemcpyAsync(out.data(), level_out_.get(), out.size() * sizeof(out[0]), cudaMemcpyDeviceToHost, stream_));
    SYLDET_CUDA(cudaMemsetAsync(level_in_.get(), 0, level_in_.size(), stream_));
    SYLDET_CUDA(cudaMemsetAsync(level_out_.get(), 0x80, level_out_.size(), stream_));
    SYLDET_CUDA(cudaStreamSynchronize(stream_));
    const double nil = std::nan("");
    for (int ch = 0; ch < n_channels_; ++ch) {
        if (input_rms) {
            double ms;
            std::memcpy(&ms, &in[ch], sizeof ms);
            input_rms[ch] = buffers_seen_ > 0 ? std::sqrt(ms) : nil;  // sqrt(meanSquareLevel), Processor.swift:168-170
        }
        if (output_max) {
            const int img = out[ch];
            const int bits = img >= 0 ? img : img ^ 0x7fffffff;
            float v;
            std::memcpy(&v, &bits, sizeof v);
            output_max[ch] = (evals_seen_ > 0 && img != (int)0x80808080) ? (double)v : nil;
        }
    }
    buffers_seen_ = 0;
    evals_seen_ = 0;
    return SYLDET_OK;
}

// ---------------------------------------------------------------------------------------------------------------
syldet_status Detector::init(const Config &cfg, int device) {
    // 409 600-byte sample ring (CSTFT.swift:61,126): at most 102 400 unconsumed floats; one append may be that long
    syldet_status st = group_.init(cfg, 1, kSampleRingFloats, device);
    if (st != SYLDET_OK) return st;
    const Config &c = group_.config();
    last_outputs_.assign(c.outputs, 0.0f);  // SyllableDetector.swift:70
    // feature ring: L*T*512 BYTES rounded up to a page (SyllableDetector.swift:63-67, TPCircularBuffer.c:49)
    const int64_t bytes = ((int64_t)c.band * c.time_range * 512 + 4095) / 4096 * 4096;
    feature_ring_cols_ = bytes / (4 * (int64_t)c.band);
    return SYLDET_OK;
}

syldet_status Detector::append(const float *samples, int64_t n) {
    if (n < 0 || (n > 0 && !samples)) return set_error(SYLDET_ERR_ARG, "bad samples");
    const Config &c = group_.config();
    // upstream consumes samples only when columns are extracted inside processNewValue (CSTFT.swift:298-302)
    const int64_t unconsumed = appended_ - cols_extracted_ * c.hop;
    if (unconsumed + n > kSampleRingFloats) return set_error(SYLDET_ERR_OVERFLOW, "Insufficient space on buffer.");
    pending_.insert(pending_.end(), samples, samples + n);
    appended_ += n;
    return SYLDET_OK;
}

int Detector::process_new_value() {
    const Config &c = group_.config();
    // "while processFourierData() {}" drains every complete column into the feature ring (SyllableDetector.swift:155)
    const int64_t cols_now = c.num_columns(appended_);
    if (cols_now - evals_returned_ > feature_ring_cols_) {
        set_error(SYLDET_ERR_OVERFLOW, "Insufficient space on buffer.");
        return -(int)SYLDET_ERR_OVERFLOW;
    }
    cols_extracted_ = cols_now;
    if (queue_head_ == queue_.size() / c.outputs) {
        queue_.clear();
        queue_head_ = 0;
        if (!pending_.empty()) {  // push everything appended since the last flush through the device
            size_t pos = 0;
            while (pos < pending_.size()) {
                const int n = (int)std::min<size_t>(pending_.size() - pos, (size_t)kSampleRingFloats);
                const float *buf = pending_.data() + pos;
                const float *outs = nullptr;
                int64_t n_new = 0;
                syldet_status st = group_.submit(&buf, n, &outs, &n_new);
                if (st != SYLDET_OK) return -(int)st;
                queue_.insert(queue_.end(), outs, outs + n_new * c.outputs);
                pos += n;
            }
            pending_.clear();
        }
    }
    if (queue_head_ == queue_.size() / c.outputs) return 0;  // fewer than L*T features buffered (SyllableDetector.swift:168-172)
    std::copy(queue_.begin() + queue_head_ * c.outputs, queue_.begin() + (queue_head_ + 1) * c.outputs, last_outputs_.begin());
    ++queue_head_;
    ++evals_returned_;
    return 1;
}

bool Detector::last_detected() const {
    const Config &c = group_.config();
    return (double)last_outputs_[0] >= c.thresholds[0];  // SyllableDetector.swift:27-31
}

int Detector::seen_syllable() {
    int ret = 0;
    for (;;) {  // SyllableDetector.swift:220-230
        const int r = process_new_value();
        if (r < 0) return r;
        if (r == 0) break;
        if (last_detected()) ret = 1;
    }
    return ret;
}

// ---------------------------------------------------------------------------------------------------------------
namespace {
// y[k] = x[b] + a (x[b+1] - x[b]), idx = offset + k*step in float32 exactly as vDSP_vramp + vDSP_vlint (Resampler.swift:52-59)
__global__ void resample_linear_kernel(const float *__restrict__ x, int64_t n_in, float offset, float step, float last, int across,
                                       float *__restrict__ y, int64_t n_out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_out; k += (int64_t)gridDim.x * blockDim.x) {
        float idx = __fadd_rn(offset, __fmul_rn((float)k, step));
        if (k == 0 && across) idx = 0.0f;
        const int64_t b = (int64_t)idx;
        const float a = __fsub_rn(idx, (float)b);
        const float x0 = x[b];
        const float x1 = b + 1 < n_in ? x[b + 1] : x[n_in - 1];  // upstream reads one past the end when up-sampling; we hold
        float v = __fadd_rn(x0, __fmul_rn(a, __fsub_rn(x1, x0)));
        if (k == 0 && across) v = __fadd_rn(__fmul_rn(last, __fsub_rn(0.0f, offset)), __fmul_rn(x[0], __fadd_rn(1.0f, offset)));
        y[k] = v;
    }
}
}  // namespace

syldet_status Resampler::init(double rate_in, double rate_out, int device) {
    if (!(rate_in > 0.0) || !(rate_out > 0.0)) return set_error(SYLDET_ERR_ARG, "sampling rates must be positive");
    step_ = (float)(rate_in / rate_out);  // Resampler.swift:32
    device_ = device;
    syldet_status st = use_device(device);
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    return SYLDET_OK;
}

Resampler::~Resampler() {
    if (stream_) {
        cudaSetDevice(device_);
        cudaStreamDestroy(stream_);
    }
}

int64_t Resampler::max_output(int64_t n_in) const {
    // offset stays within (-step, 1]; one extra for rounding
    return (int64_t)(((double)n_in + (double)step_ + 1.0) / (double)step_) + 2;
}

syldet_status Resampler::process(const float *in, int64_t n_in, float *out, int64_t cap, int64_t *n_out) {
    *n_out = 0;
    if (n_in < 0 || (n_in > 0 && !in)) return set_error(SYLDET_ERR_ARG, "bad input");
    if (n_in == 0) return SYLDET_OK;
    const bool across = offset_ < 0;                                          // Resampler.swift:37
    const int64_t n = (int64_t)(((float)n_in - offset_) / step_);              // Resampler.swift:40
    if (n > cap) return set_error(SYLDET_ERR_ARG, "output capacity too small");
    if (n <= 0) {  // upstream indexes indices[-1] here; we carry the phase forward instead
        offset_ = offset_ - (float)n_in;
        last_ = in[n_in - 1];
        return SYLDET_OK;
    }
    syldet_status st = use_device(device_);
    if (st != SYLDET_OK) return st;
    st = d_in_.reserve((size_t)n_in * sizeof(float));
    if (st != SYLDET_OK) return st;
    st = d_out_.reserve((size_t)n * sizeof(float));
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaMemcpyAsync(d_in_.get(), in, (size_t)n_in * sizeof(float), cudaMemcpyHostToDevice, stream_));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    resample_linear_kernel<<<blocks, 256, 0, stream_>>>(d_in_.as<float>(), n_in, offset_, step_, last_, across ? 1 : 0,
                                                         d_out_.as<float>(), n);
    SYLDET_CUDA(cudaGetLastError());
    SYLDET_CUDA(cudaMemcpyAsync(out, d_out_.get(), (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, stream_));
    SYLDET_CUDA(cudaStreamSynchronize(stream_));
    // state carried to the next buffer (Resampler.swift:65-66); float32, multiply and add unfused like the kernel
    volatile float ramp = (float)(n - 1) * step_;
    volatile float last_idx = offset_ + ramp;
    if (n == 1 && across) last_idx = 0.0f;
    volatile float t = last_idx + step_;
    offset_ = t - (float)(n_in - 1);
    last_ = in[n_in - 1];
    *n_out = n;
    return SYLDET_OK;
}

}  // namespace syldet
