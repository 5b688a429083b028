// ResamplerLinear.resampleVector (Common/Resampler.swift:35-70), one output sample, in the reference's float32 operation order:
//   indices[k] = offset + k * step            (vDSP_vramp, :52; multiply and add unfused)
//   indices[0] = 0 when offset < 0            (:54-56)
//   y[k] = x[b] + a (x[b+1] - x[b])           (vDSP_vlint, :59; b = trunc(index), a = its fraction)
//   y[0] = last (0 - offset) + x[0] (1 + offset) when offset < 0   (:61-63)
// Deviation: upstream reads x[n_in] - one past the buffer - when up-sampling; we hold x[n_in - 1].
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace syldet {

__device__ __forceinline__ float resample_linear_point(const float *__restrict__ x, int64_t n_in, float offset, float step, float last,
                                                       bool across, int64_t k) {
    float idx = __fadd_rn(offset, __fmul_rn((float)k, step));
    if (k == 0 && across) idx = 0.0f;
    const int64_t b = (int64_t)idx;
    const float a = __fsub_rn(idx, (float)b);
    const float x0 = x[b];
    const float x1 = b + 1 < n_in ? x[b + 1] : x[n_in - 1];
    float v = __fadd_rn(x0, __fmul_rn(a, __fsub_rn(x1, x0)));
    if (k == 0 && across) v = __fadd_rn(__fmul_rn(last, __fsub_rn(0.0f, offset)), __fmul_rn(x[0], __fadd_rn(1.0f, offset)));
    return v;
}

// Host side of the same object: the data-independent part of the state (`offset`) and the output count of the next buffer.
struct LinearResamplerPhase {
    float step = 1.0f, offset = 0.0f;   // Resampler.swift:24-26
    // numSamplesOut of the next buffer of n_in samples (:40); <= 0 where upstream would index indices[-1]
    int64_t plan(int64_t n_in) const { return (int64_t)(((float)n_in - offset) / step); }
    // the carry to the next buffer (:65), float32, multiply and add unfused like the kernel
    void advance(int64_t n_in, int64_t n_out) {
        if (n_out <= 0) {   // upstream crashes here; we carry the phase forward
            offset = offset - (float)n_in;
            return;
        }
        const bool across = offset < 0;
        volatile float ramp = (float)(n_out - 1) * step;
        volatile float last_idx = offset + ramp;
        if (n_out == 1 && across) last_idx = 0.0f;
        volatile float t = last_idx + step;
        offset = t - (float)(n_in - 1);
    }
};

}  // namespace syldet
