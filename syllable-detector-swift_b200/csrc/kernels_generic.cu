// Reference-order CUDA path: handles every configuration the text format can express.
//
// This file is compiled with -fmad=false and performs the float32 operations in the order the Swift code calls vDSP
// (window multiply -> zero pad -> even/odd packed radix-2 real FFT -> sqrt(re^2+im^2)/2 -> band slice -> [scaling] ->
// processing chain -> left-to-right W.x + b -> transfer -> reverse output map -> double(out) >= threshold), so its
// results differ from a scalar CPU evaluation only through the device's tanhf/expf/logf/log10f.
//   STFT column:  Common/CircularShortTimeFourierTransform.swift:280-337 (extractPower)
//   feature ring: Common/SyllableDetector.swift:134-217
//   network:      Common/NeuralNet.swift:41-228, 294-326, 366-377
//   decision:     SyllableDetectorCLI/TrackDetector.swift:71-77, Common/SyllableDetector.swift:27-31
// The fused kernel (kernels_fused.cu) is the fast path; this one is the general one and the on-device cross-check.
#include "kernels.hpp"

namespace syldet {

namespace {

constexpr int kWarp = 32;

__device__ __forceinline__ float scale_value(float v, int scaling) {
    if (scaling == SYLDET_SCALING_DB) return 20.0f * log10f(v / 1.0f);  // vDSP_vdbcon, amplitude flag (SyllableDetector.swift:195)
    if (scaling == SYLDET_SCALING_LOG) return logf(v);                   // intent of SyllableDetector.swift:207 (upstream call is broken)
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// PCM ingest: int16 and/or interleaved -> planar float32 (AVAssetReader's LPCM Float32 non-interleaved conversion,
// SyllableDetector.swift:19-23; appendInterleavedData, CSTFT.swift:203-217).
__global__ void ingest_kernel(const void *__restrict__ src, int format, int interleaved, int n_channels, int64_t n_samples,
                              int64_t src_stride, float *__restrict__ dst, int64_t dst_stride) {
    const int64_t total = (int64_t)n_channels * n_samples;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int ch;
        int64_t s;
        if (interleaved) { s = i / n_channels; ch = (int)(i - s * n_channels); }  // consecutive threads read consecutive source
        else { ch = (int)(i / n_samples); s = i - (int64_t)ch * n_samples; }
        const int64_t si = interleaved ? i : (int64_t)ch * src_stride + s;
        float v = format == SYLDET_PCM_S16 ? (float)((const int16_t *)src)[si] * (1.0f / 32768.0f) : ((const float *)src)[si];
        dst[(int64_t)ch * dst_stride + s] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per STFT column, iterative radix-2 DIT over M = N/2 complex points held in shared memory.
__global__ void stft_band_generic_kernel(const DevNet *__restrict__ netp, const float *__restrict__ pcm, int64_t ch_stride,
                                         int64_t col0, int64_t n_cols, float *__restrict__ feat) {
    extern __shared__ float smem[];
    const DevNet &net = *netp;
    const int N = net.fft_len, M = N / 2, W = net.win_len, L = net.band;
    const int warps = blockDim.x / kWarp, warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    float *zr = smem + (size_t)warp * N, *zi = zr + M;
    const int ch = blockIdx.y;
    const float *x_ch = pcm + (int64_t)ch * ch_stride;
    float *feat_ch = feat + (int64_t)ch * n_cols * L;
    int bits = 0;
    while ((1 << bits) < M) ++bits;

    for (int64_t c = (int64_t)blockIdx.x * warps + warp; c < n_cols; c += (int64_t)gridDim.x * warps) {
        const float *fr = x_ch + (col0 + c) * net.hop + net.gap;
        if (M == 1) {
            if (lane == 0) {
                float a = fr[0] * net.window[0], b = W > 1 ? fr[1] * net.window[1] : 0.0f;
                feat_ch[c * L] = scale_value(fabsf(2.0f * (a + b)) / 2.0f, net.scaling);
            }
            continue;
        }
        for (int n = lane; n < M; n += kWarp) {
            const int r = (int)(__brev((unsigned)n) >> (32 - bits));
            const int m0 = 2 * n, m1 = 2 * n + 1;
            zr[r] = m0 < W ? fr[m0] * net.window[m0] : 0.0f;
            zi[r] = m1 < W ? fr[m1] * net.window[m1] : 0.0f;
        }
        __syncwarp();
        for (int len = 2; len <= M; len <<= 1) {
            const int half = len >> 1, step = N / len;
            for (int idx = lane; idx < (M >> 1); idx += kWarp) {
                const int j = idx % half, a = (idx / half) * len + j, b = a + half;
                const float2 w = net.twiddle[j * step];
                const float tr = zr[b] * w.x - zi[b] * w.y;
                const float ti = zr[b] * w.y + zi[b] * w.x;
                const float ar = zr[a], ai = zi[a];
                zr[b] = ar - tr;
                zi[b] = ai - ti;
                zr[a] = ar + tr;
                zi[a] = ai + ti;
            }
            __syncwarp();
        }
        for (int f = lane; f < L; f += kWarp) {
            const int k = net.k0 + f;
            float mag;
            if (k == 0) {
                const float re0 = 2.0f * (zr[0] + zi[0]);
                mag = sqrtf(re0 * re0 + 0.0f * 0.0f) / 2.0f;
            } else {
                const float ar = zr[k], ai = zi[k], br = zr[M - k], bi = -zi[M - k];
                const float sr = ar + br, si = ai + bi, dr = ar - br, di = ai - bi;
                const float2 w = net.twiddle[k];
                const float re = sr + (w.x * di + w.y * dr);
                const float im = si - (w.x * dr - w.y * di);
                mag = sqrtf(re * re + im * im) / 2.0f;
            }
            feat_ch[c * L + f] = scale_value(mag, net.scaling);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per evaluation: literal processing chain, lanes own output neurons (each sum is left-to-right in one lane).
__device__ void apply_input_processing(const DevProcessing &p, float *x, int n, int lane) {
    switch (p.function) {
        case SYLDET_PROC_MAPMINMAX:
            for (int i = lane; i < n; i += kWarp) { float t = (x[i] - p.xoff[i]) * p.gain[i]; x[i] = t + p.y; }
            break;
        case SYLDET_PROC_MAPSTD:
            for (int i = lane; i < n; i += kWarp) {
                float t = (x[i] - p.xoff[i]) * p.gain[i];
                x[i] = (0 != p.y) ? t + p.y : t;
            }
            break;
        case SYLDET_PROC_L2NORMALIZE: {
            float d = 0.0f;
            if (lane == 0) {
                float ss = 0.0f;
                for (int i = 0; i < n; ++i) ss += x[i] * x[i];
                d = sqrtf(ss);
            }
            d = __shfl_sync(0xffffffffu, d, 0);
            for (int i = lane; i < n; i += kWarp) x[i] = x[i] / d;
            break;
        }
        case SYLDET_PROC_NORMALIZE: {
            float mn = 0.0f, mx = 0.0f;
            if (lane == 0) {
                mn = mx = x[0];
                for (int i = 1; i < n; ++i) { if (x[i] < mn) mn = x[i]; if (x[i] > mx) mx = x[i]; }
            }
            mn = __shfl_sync(0xffffffffu, mn, 0);
            mx = __shfl_sync(0xffffffffu, mx, 0);
            const float range = mx - mn;
            if (0 == range) {
                for (int i = lane; i < n; i += kWarp) x[i] = -1.0f;
            } else {
                const float slope = 2.0f / range, icpt = (0 - mn - mx) / range;
                for (int i = lane; i < n; i += kWarp) { float t = x[i] * slope; x[i] = t + icpt; }
            }
            break;
        }
        case SYLDET_PROC_NORMALIZESTD: {
            float mean = 0.0f, sd = 0.0f;
            if (lane == 0) {
                float s = 0.0f, v = 0.0f;
                for (int i = 0; i < n; ++i) s += x[i];
                mean = s / (float)n;
                for (int i = 0; i < n; ++i) { float d = x[i] - mean; v += d * d; }
                sd = sqrtf(v / (float)n);
            }
            mean = __shfl_sync(0xffffffffu, mean, 0);
            sd = __shfl_sync(0xffffffffu, sd, 0);
            for (int i = lane; i < n; i += kWarp) x[i] = (x[i] - mean) / sd;
            break;
        }
        default: break;
    }
}

__device__ __forceinline__ float apply_transfer(int tf, float v) {
    switch (tf) {
        case SYLDET_TF_TANSIG: return tanhf(v);
        case SYLDET_TF_LOGSIG: { float t = v * -1.0f; t = expf(t); t = t + 1.0f; return 1.0f / t; }
        case SYLDET_TF_SATLIN: return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        default: return v;
    }
}

__global__ void nn_generic_kernel(const DevNet *__restrict__ netp, const float *__restrict__ feat, int64_t n_cols,
                                  int64_t n_evals, int64_t eval0, int64_t evals_total, int detect_rule,
                                  float *__restrict__ all_out, EventSink sink) {
    extern __shared__ float smem[];
    const DevNet &net = *netp;
    const int warps = blockDim.x / kWarp, warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int I = net.inputs, O = net.outputs, L = net.band;
    float *buf0 = smem + (size_t)warp * 2 * net.max_width, *buf1 = buf0 + net.max_width;
    const int ch = blockIdx.y;
    const float *feat_ch = feat + (int64_t)ch * n_cols * L;

    for (int64_t j = (int64_t)blockIdx.x * warps + warp; j < n_evals; j += (int64_t)gridDim.x * warps) {
        float *cur = buf0, *nxt = buf1;
        for (int i = lane; i < I; i += kWarp) cur[i] = feat_ch[j * L + i];  // v[t*L+f]: contiguous in the column stream
        __syncwarp();
        for (int k = 0; k < net.n_ip; ++k) {
            apply_input_processing(net.ip[k], cur, I, lane);
            __syncwarp();
        }
        for (int l = 0; l < net.n_layers; ++l) {
            const DevLayer &ly = net.layers[l];
            for (int o = lane; o < ly.outputs; o += kWarp) {
                const float *w = ly.w + (size_t)o * ly.inputs;
                float acc = 0.0f;
                for (int i = 0; i < ly.inputs; ++i) acc += w[i] * cur[i];
                nxt[o] = apply_transfer(ly.transfer, acc + ly.b[o]);
            }
            __syncwarp();
            float *t = cur; cur = nxt; nxt = t;
        }
        bool hit = false;
        for (int o = lane; o < O; o += kWarp) {
            float v = cur[o];
            for (int k = 0; k < net.n_op; ++k) {  // reverse transforms, index order (NeuralNet.swift:316-323)
                const DevProcessing &p = net.op[k];
                float t = v + (0 - p.y);
                t = t / p.gain[o];
                v = t + p.xoff[o];
            }
            cur[o] = v;
            const bool over = (double)v >= net.thresholds[o];  // NaN compares false
            if (detect_rule == SYLDET_DETECT_FIRST_OUTPUT ? (o == 0 && over) : over) hit = true;
            if (all_out) all_out[((int64_t)ch * evals_total + eval0 + j) * O + o] = v;
        }
        const bool any = __any_sync(0xffffffffu, hit);
        __syncwarp();
        if (any && lane == 0) sink_push(sink, ch, eval0 + j, cur, O);
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_ingest(const void *src, int format, int interleaved, int n_channels, int64_t n_samples, int64_t src_stride,
                          float *dst, int64_t dst_stride, cudaStream_t stream) {
    const int64_t total = (int64_t)n_channels * n_samples;
    if (total <= 0) return cudaSuccess;
    int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    ingest_kernel<<<blocks, 256, 0, stream>>>(src, format, interleaved, n_channels, n_samples, src_stride, dst, dst_stride);
    return cudaGetLastError();
}

cudaError_t launch_stft_band_generic(const DevNet *d_net, int fft_len, const float *pcm, int64_t ch_stride, int n_channels,
                                     int64_t col0, int64_t n_cols, float *feat, cudaStream_t stream) {
    if (n_cols <= 0 || n_channels <= 0) return cudaSuccess;
    int warps = 8;
    while (warps > 1 && (size_t)warps * fft_len * sizeof(float) > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * fft_len * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(stft_band_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t bx = (n_cols + warps - 1) / warps;
    const int64_t cap = std::max<int64_t>(1, (148 * 8) / n_channels);
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)n_channels);
    stft_band_generic_kernel<<<grid, warps * 32, smem, stream>>>(d_net, pcm, ch_stride, col0, n_cols, feat);
    return cudaGetLastError();
}

cudaError_t launch_nn_generic(const DevNet *d_net, int max_width, const float *feat, int n_channels, int64_t n_cols,
                              int64_t n_evals, int64_t eval0, int64_t evals_total, int detect_rule, float *all_out,
                              EventSink sink, cudaStream_t stream) {
    if (n_evals <= 0 || n_channels <= 0) return cudaSuccess;
    int warps = 8;
    while (warps > 1 && (size_t)warps * 2 * max_width * sizeof(float) > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * 2 * max_width * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(nn_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t bx = (n_evals + warps - 1) / warps;
    const int64_t cap = std::max<int64_t>(1, (148 * 8) / n_channels);
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)n_channels);
    nn_generic_kernel<<<grid, warps * 32, smem, stream>>>(d_net, feat, n_cols, n_evals, eval0, evals_total, detect_rule, all_out, sink);
    return cudaGetLastError();
}

}  // namespace syldet
