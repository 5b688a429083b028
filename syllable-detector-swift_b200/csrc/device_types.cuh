// Device-visible description of one detector configuration, shared by all kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace syldet {

constexpr int kMaxProcessing = 8;
constexpr int kMaxLayers = 8;

struct DevProcessing {
    int function;       // SYLDET_PROC_*
    float y;            // yMin / yMean
    const float *xoff;  // device, [n] (mapminmax / mapstd only)
    const float *gain;
};

struct DevLayer {
    int inputs, outputs, transfer;
    const float *w;  // device, row-major [outputs][inputs]
    const float *b;
};

// Geometry + literal network (reference operation order). Lives in device global memory.
struct DevNet {
    int fft_len, win_len, gap, hop, k0, band, time_range, inputs, outputs, scaling;
    int n_ip, n_op, n_layers, max_width;
    DevProcessing ip[kMaxProcessing];
    DevProcessing op[kMaxProcessing];
    DevLayer layers[kMaxLayers];
    const double *thresholds;  // [outputs]
    const float *window;       // [win_len]  Hamming, N-denominator (CSTFT.swift:24, SyllableDetector.swift:43)
    const float2 *twiddle;     // [fft_len/2] e^{-2 pi i k / fft_len}
};

// One raw detection (before debounce): evaluation index j of a channel. Host turns j into the sample number.
struct DevEvent {
    int32_t channel;
    int32_t reserved;
    int64_t eval;
};

struct EventSink {
    unsigned long long *count;  // total detections seen (may exceed capacity)
    DevEvent *events;           // [capacity]
    float *outputs;             // [capacity][n_outputs]
    unsigned long long capacity;
};

__device__ __forceinline__ void sink_push(const EventSink &s, int channel, int64_t eval, const float *out, int n_out) {
    unsigned long long slot = atomicAdd(s.count, 1ULL);
    if (slot < s.capacity) {
        s.events[slot] = DevEvent{channel, 0, eval};
        for (int o = 0; o < n_out; ++o) s.outputs[slot * n_out + o] = out[o];
    }
}

}  // namespace syldet
