// Streaming front-ends on top of the batch kernels:
//   StreamGroup  - Processor.swift's shape: many live channels, one small buffer per channel per tick
//                  (SyllableDetector/Processor.swift:102-149, AudioInterface.swift:474 frameSize 32)
//   Detector     - class SyllableDetector, one stream with appendAudioData / processNewValue / lastOutputs
//                  (Common/SyllableDetector.swift:13-231)
//   Resampler    - ResamplerLinear (Common/Resampler.swift:20-70), the index ramp and interpolation run in a kernel
// Audio lives in a device-resident per-channel buffer; each tick uploads only the new samples, launches the detection
// kernel over the unconsumed tail of every channel at once and reads back the few new network outputs.
#include "stream.hpp"

#include <algorithm>
#include <cstring>

namespace syldet {

// ---------------------------------------------------------------------------------------------------------------
syldet_status StreamGroup::init(const Config &cfg, int n_channels, int max_buffer, int device) {
    if (n_channels <= 0 || n_channels > 65535) return set_error(SYLDET_ERR_ARG, "n_channels out of range");
    if (max_buffer <= 0) return set_error(SYLDET_ERR_ARG, "max_buffer must be positive");
    syldet_status st = batch_.init(cfg, device);
    if (st != SYLDET_OK) return st;
    n_channels_ = n_channels;
    max_buffer_ = max_buffer;
    const Config &c = batch_.model().config();
    // room for the retained tail (one full feature window) plus many ticks before a compaction is needed
    const int64_t tail = c.samples_for_evals(1) + c.hop;
    const int64_t ticks = std::min<int64_t>(64, std::max<int64_t>(2, 262144 / max_buffer));
    cap_ = ((tail + ticks * max_buffer + 4095) / 4096) * 4096;
    st = ring_[0].reserve((size_t)n_channels * cap_ * sizeof(float));
    if (st != SYLDET_OK) return st;
    st = ring_[1].reserve((size_t)n_channels * cap_ * sizeof(float));
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    SYLDET_CUDA(cudaMallocHost(&h_in_, (size_t)n_channels * max_buffer * sizeof(float)));
    max_new_ = c.num_evals(tail + max_buffer) + 2;
    SYLDET_CUDA(cudaMallocHost(&h_out_, (size_t)n_channels * max_new_ * c.outputs * sizeof(float)));
    st = d_out_.reserve((size_t)n_channels * max_new_ * c.outputs * sizeof(float));
    return st;
}

StreamGroup::~StreamGroup() {
    if (stream_) {
        cudaSetDevice(batch_.model().device());
        cudaStreamSynchronize(stream_);
        cudaStreamDestroy(stream_);
    }
    if (h_in_) cudaFreeHost(h_in_);
    if (h_out_) cudaFreeHost(h_out_);
}

syldet_status StreamGroup::submit(const float *const *bufs, int n, const float **outs, int64_t *n_new) {
    *n_new = 0;
    *outs = h_out_;
    if (n < 0 || n > max_buffer_) return set_error(SYLDET_ERR_ARG, "buffer longer than max_buffer");
    if (n == 0) return SYLDET_OK;
    syldet_status st = use_device(batch_.model().device());
    if (st != SYLDET_OK) return st;
    const Config &c = batch_.model().config();
    // compaction: keep only samples from the start of the next evaluation's window
    if (fill_ + n > cap_) {
        const int64_t keep_from = next_eval_ * c.hop - base_;
        const int64_t keep = fill_ - keep_from;
        if (keep + n > cap_) return set_error(SYLDET_ERR_OVERFLOW, "Insufficient space on buffer.");
        SYLDET_CUDA(cudaMemcpy2DAsync(ring_[cur_ ^ 1].get(), cap_ * sizeof(float), ring_[cur_].as<float>() + keep_from,
                                      cap_ * sizeof(float), keep * sizeof(float), n_channels_, cudaMemcpyDeviceToDevice, stream_));
        cur_ ^= 1;
        base_ += keep_from;
        fill_ = keep;
    }
    for (int ch = 0; ch < n_channels_; ++ch) std::memcpy(h_in_ + (size_t)ch * n, bufs[ch], (size_t)n * sizeof(float));
    SYLDET_CUDA(cudaMemcpy2DAsync(ring_[cur_].as<float>() + fill_, cap_ * sizeof(float), h_in_, (size_t)n * sizeof(float),
                                  (size_t)n * sizeof(float), n_channels_, cudaMemcpyHostToDevice, stream_));
    fill_ += n;
    total_ += n;
    const int64_t avail = c.num_evals(total_) - next_eval_;
    if (avail <= 0) {
        SYLDET_CUDA(cudaStreamSynchronize(stream_));  // h_in_ is reused by the next tick
        return SYLDET_OK;
    }
    if (avail > max_new_) return set_error(SYLDET_ERR_OVERFLOW, "more evaluations pending than the stream was sized for");
    const int64_t seg0 = next_eval_ * c.hop - base_;
    st = batch_.launch_device(ring_[cur_].as<float>() + seg0, n_channels_, fill_ - seg0, cap_, SYLDET_LAYOUT_PLANAR,
                              SYLDET_DETECT_FIRST_OUTPUT, d_out_.as<float>(), stream_);
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaMemcpyAsync(h_out_, d_out_.get(), (size_t)n_channels_ * avail * c.outputs * sizeof(float),
                                cudaMemcpyDeviceToHost, stream_));
    SYLDET_CUDA(cudaStreamSynchronize(stream_));
    next_eval_ += avail;
    *n_new = avail;
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
