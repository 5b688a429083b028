// Epilogue shared by the fused kernels: one thread = one evaluation, reading its T band columns from the shared-memory
// ring (NeuralNet.apply, Common/NeuralNet.swift:294-326; threshold test, SyllableDetectorCLI/TrackDetector.swift:71-77).
#pragma once
#include "kernels.hpp"

namespace syldet {
namespace {

__device__ __forceinline__ float scale_value(float v, int scaling) {
    if (scaling == SYLDET_SCALING_DB) return 20.0f * log10f(v);
    if (scaling == SYLDET_SCALING_LOG) return logf(v);
    return v;
}

__device__ __forceinline__ float transfer(int tf, float v) {
    switch (tf) {
        case SYLDET_TF_TANSIG: return tanhf(v);
        case SYLDET_TF_LOGSIG: return 1.0f / (1.0f + expf(-v));
        case SYLDET_TF_SATLIN: return fminf(fmaxf(v, 0.0f), 1.0f);
        default: return v;
    }
}

// Out-of-line copy for the warp-specialised kernel, whose roles share the instruction cache: one body instead of one per use.
__device__ __noinline__ float scale_value_nl(float v, int scaling) { return scale_value(v, scaling); }

__device__ __forceinline__ float ex2_fast(float x) {  // MUFU.EX2, max relative error 2^-22
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_fast(float x) {  // MUFU.RCP, 1 ulp
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Transfer functions on the special-function unit (tensor-core kernel): tanh(v) = 1 - 2 / (e^{2v} + 1), absolute error
// <= 4e-7 for every v (the 1e-5 output tolerance leaves room for it); e^{2v} = inf gives 1, 0 gives -1, NaN stays NaN.
__device__ __forceinline__ float transfer_fast(int tf, float v) {
    if (tf == SYLDET_TF_TANSIG) return fmaf(-2.0f, rcp_fast(ex2_fast(v * 2.8853900817779268f) + 1.0f), 1.0f);
    if (tf == SYLDET_TF_LOGSIG) return rcp_fast(1.0f + ex2_fast(v * -1.4426950408889634f));
    if (tf == SYLDET_TF_SATLIN) return fminf(fmaxf(v, 0.0f), 1.0f);
    return v;
}

__device__ __forceinline__ float sqrt_fast(float x) {  // MUFU.SQRT, max relative error 2^-23
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float sqrt_fast_ftz(float x) {  // as sqrt_fast, denormal inputs read as zero (no scaling fix-up code)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- epilogue: one thread = one evaluation -----------------------------------------------------------------------
template <int HP, int STAT>
__device__ __forceinline__ void gather_layer0(const FusedParams &p, const float *ring, int slot, float (&acc)[HP], float &s0,
                                              float &s1) {
    // s0/s1: STAT_L2 -> (sum x^2, -) ; STAT_MINMAX -> (min, max) ; STAT_STD -> (sum x, -)
    const int L = p.band, T = p.time_range;
    int widx = 0;
    for (int t = 0; t < T; ++t) {
        const float *row = ring + slot * p.band_pitch;
#pragma unroll 4
        for (int f = 0; f < L; ++f) {
            const float x = row[f];
            if constexpr (STAT == FUSED_STAT_L2) s0 = fmaf(x, x, s0);
            if constexpr (STAT == FUSED_STAT_MINMAX) { s0 = fminf(s0, x); s1 = fmaxf(s1, x); }
            if constexpr (STAT == FUSED_STAT_STD) s0 += x;
            const float4 wa = *reinterpret_cast<const float4 *>(&p.w0[widx]);
            acc[0] = fmaf(x, wa.x, acc[0]);
            acc[1] = fmaf(x, wa.y, acc[1]);
            acc[2] = fmaf(x, wa.z, acc[2]);
            acc[3] = fmaf(x, wa.w, acc[3]);
            if constexpr (HP == 8) {
                const float4 wb = *reinterpret_cast<const float4 *>(&p.w0[widx + 4]);
                acc[4] = fmaf(x, wb.x, acc[4]);
                acc[5] = fmaf(x, wb.y, acc[5]);
                acc[6] = fmaf(x, wb.z, acc[6]);
                acc[7] = fmaf(x, wb.w, acc[7]);
            }
            widx += HP;
        }
        if (++slot == p.ring_cols) slot = 0;
    }
}

// Everything after the layer-0 dot products: z_h = acc_h / alpha_div + beta V_h + B'_h, transfer functions, remaining
// layers, reverse output maps, threshold test. Returns true when the evaluation counts as a detection.
template <int HP>
__device__ __forceinline__ bool finish_eval(const FusedParams &p, int detect_rule, const float (&acc)[HP], float alpha_div, float beta,
                                            bool constant_input, float (&out)[kFusedMaxOut]) {
    float a[kFusedMaxHidden], b[kFusedMaxHidden];
#pragma unroll
    for (int h = 0; h < kFusedMaxHidden; ++h) {
        float z = 0.0f;
        if (h < HP) {
            const float u = constant_input ? 0.0f : acc[h] / alpha_div;
            z = u + fmaf(beta, p.v[h], p.bprime[h]);
            z = transfer(p.tf[0], z);
        }
        a[h] = z;
    }
    for (int l = 1; l < p.n_layers; ++l) {
        const float *w = &p.rest_w[(l - 1) * kFusedMaxHidden * kFusedMaxHidden];
        const float *bias = &p.rest_b[(l - 1) * kFusedMaxHidden];
#pragma unroll
        for (int o = 0; o < kFusedMaxHidden; ++o) {
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < kFusedMaxHidden; ++i) s = fmaf(w[o * kFusedMaxHidden + i], a[i], s);
            b[o] = transfer(p.tf[l], s + bias[o]);
        }
#pragma unroll
        for (int o = 0; o < kFusedMaxHidden; ++o) a[o] = b[o];
    }
    bool hit = false;
#pragma unroll
    for (int o = 0; o < kFusedMaxOut; ++o) {
        float v = a[o];
        if (o < p.n_out) {
            for (int k = 0; k < p.n_op; ++k) {  // reverse maps in index order (NeuralNet.swift:316-323)
                v = (v + (0 - p.op_y[k])) / p.op_gain[k * kFusedMaxOut + o] + p.op_xoff[k * kFusedMaxOut + o];
            }
            const bool over = (double)v >= p.thr[o];  // TrackDetector.swift:72; NaN -> false
            if (over && (detect_rule == SYLDET_DETECT_ANY_OUTPUT || o == 0)) hit = true;
        }
        out[o] = v;
    }
    return hit;
}

// a[k] / a[k] = v with a runtime index, without spilling the array to local memory
template <int N>
__device__ __forceinline__ float pick(const float (&a)[N], int k) {
    float v = a[0];
#pragma unroll
    for (int i = 1; i < N; ++i) v = (k == i) ? a[i] : v;
    return v;
}
template <int N>
__device__ __forceinline__ void put(float (&a)[N], int k, float v) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = (k == i) ? v : a[i];
}

// Compact form of the network tail (layers >= 1, reverse output maps, threshold test) for callers that hold the layer-0
// activations in a[] (zero padded): loops run over the real layer widths, so a 4 -> 1 tail costs ~40 instructions.
__device__ __forceinline__ bool network_tail(const FusedParams &p, int detect_rule, float (&a)[kFusedMaxHidden], float (&out)[kFusedMaxOut]) {
    for (int l = 1; l < p.n_layers; ++l) {
        const float *w = &p.rest_w[(l - 1) * kFusedMaxHidden * kFusedMaxHidden];
        const float *bias = &p.rest_b[(l - 1) * kFusedMaxHidden];
        float b[kFusedMaxHidden];
#pragma unroll
        for (int o = 0; o < kFusedMaxHidden; ++o) b[o] = 0.0f;
        const int wo = p.width[l];
#pragma unroll 1
        for (int o = 0; o < wo; ++o) {
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < kFusedMaxHidden; ++i) s = fmaf(w[o * kFusedMaxHidden + i], a[i], s);  // padding weights are zero
            put(b, o, transfer_fast(p.tf[l], s + bias[o]));
        }
#pragma unroll
        for (int o = 0; o < kFusedMaxHidden; ++o) a[o] = b[o];
    }
    bool hit = false;
#pragma unroll
    for (int o = 0; o < kFusedMaxOut; ++o) out[o] = 0.0f;
#pragma unroll 1
    for (int o = 0; o < p.n_out; ++o) {
        float v = pick(a, o);
        for (int k = 0; k < p.n_op; ++k)  // reverse maps in index order (NeuralNet.swift:316-323)
            v = (v + (0 - p.op_y[k])) / p.op_gain[k * kFusedMaxOut + o] + p.op_xoff[k * kFusedMaxOut + o];
        const bool over = v >= p.thr_f[o];  // == (double)v >= thr (TrackDetector.swift:72); NaN -> false
        if (over && (detect_rule == SYLDET_DETECT_ANY_OUTPUT || o == 0)) hit = true;
        put(out, o, v);
    }
    return hit;
}

// Window statistic -> (alpha_div, beta, constant_input) for the two-moment-free normalisers.
__device__ __forceinline__ void stat_to_affine(int window_stat, float s0, float s1, float &alpha_div, float &beta, bool &constant_input) {
    alpha_div = 1.0f;
    beta = 0.0f;
    constant_input = false;
    if (window_stat == FUSED_STAT_L2) {            // x / sqrt(sum x^2)  (NeuralNet.swift:47-59)
        alpha_div = sqrtf(s0);
    } else if (window_stat == FUSED_STAT_MINMAX) {  // x * 2/range + (-mn-mx)/range  (NeuralNet.swift:69-96)
        const float range = s1 - s0;
        if (0 == range) { constant_input = true; beta = -1.0f; }
        else { alpha_div = range * 0.5f; beta = (0 - s0 - s1) / range; }
    }
}

template <int HP>
__device__ __forceinline__ bool evaluate(const FusedParams &p, int detect_rule, const float *ring, int slot, float (&out)[kFusedMaxOut]) {
    float acc[HP];
#pragma unroll
    for (int h = 0; h < HP; ++h) acc[h] = 0.0f;
    float alpha_div = 1.0f, beta = 0.0f;  // z = acc / alpha_div + beta * V + B'
    bool constant_input = false;          // `normalize` of a flat window: every input becomes -1 (NeuralNet.swift:84-88)
    switch (p.window_stat) {
        case FUSED_STAT_L2: {  // x / sqrt(sum x^2)  (NeuralNet.swift:47-59)
            float ss = 0.0f, unused = 0.0f;
            gather_layer0<HP, FUSED_STAT_L2>(p, ring, slot, acc, ss, unused);
            alpha_div = sqrtf(ss);
            break;
        }
        case FUSED_STAT_MINMAX: {  // x * 2/range + (-mn-mx)/range  (NeuralNet.swift:69-96)
            float mn = INFINITY, mx = -INFINITY;
            gather_layer0<HP, FUSED_STAT_MINMAX>(p, ring, slot, acc, mn, mx);
            const float range = mx - mn;
            if (0 == range) { constant_input = true; beta = -1.0f; }
            else { alpha_div = range * 0.5f; beta = (0 - mn - mx) / range; }
            break;
        }
        case FUSED_STAT_STD: {  // (x - mean) / std_pop  (NeuralNet.swift:105-108)
            float sum = 0.0f, unused = 0.0f;
            gather_layer0<HP, FUSED_STAT_STD>(p, ring, slot, acc, sum, unused);
            const int n = p.band * p.time_range;
            const float mean = sum / (float)n;
            float var = 0.0f;
            int s = slot;
            for (int t = 0; t < p.time_range; ++t) {
                const float *row = ring + s * p.band_pitch;
                for (int f = 0; f < p.band; ++f) { const float d = row[f] - mean; var = fmaf(d, d, var); }
                if (++s == p.ring_cols) s = 0;
            }
            alpha_div = sqrtf(var / (float)n);
            beta = -mean / alpha_div;
            break;
        }
        default: {
            float a = 0.0f, b = 0.0f;
            gather_layer0<HP, FUSED_STAT_NONE>(p, ring, slot, acc, a, b);
        }
    }
    return finish_eval<HP>(p, detect_rule, acc, alpha_div, beta, constant_input, out);
}

}  // namespace
}  // namespace syldet
