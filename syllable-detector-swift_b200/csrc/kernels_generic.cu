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
#include "ptx_sm100.cuh"
#include "resample.cuh"

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
        float v;
        if (format == SYLDET_PCM_S16) v = (float)((const int16_t *)src)[si] * (1.0f / 32768.0f);
        else if (format == SYLDET_PCM_S24) {   // packed little-endian 24-bit (WAV): x / 2^23
            const unsigned char *b = (const unsigned char *)src + 3 * si;
            v = (float)((int)b[0] | ((int)b[1] << 8) | ((int)(signed char)b[2] << 16)) * (1.0f / 8388608.0f);
        } else v = ((const float *)src)[si];
        dst[(int64_t)ch * dst_stride + s] = v;
    }
}

// Planar int16 -> planar float32, the shape of a WAV corpus uploaded channel by channel: one channel per blockIdx.y (no index
// divisions), eight samples per thread: one 128-bit load, two 128-bit stores. Pure HBM stream (2 B in + 4 B out per sample).
__global__ void __launch_bounds__(256) ingest_planar_s16_kernel(const int16_t *__restrict__ src, int64_t n_samples, int64_t src_stride,
                                                                float *__restrict__ dst, int64_t dst_stride) {
    const int16_t *x = src + (int64_t)blockIdx.y * src_stride;
    float *y = dst + (int64_t)blockIdx.y * dst_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const int64_t n8 = aligned ? n_samples / 8 : 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n8; v += (int64_t)gridDim.x * blockDim.x) {
        const int4 q = __ldg(reinterpret_cast<const int4 *>(x) + v);
        const int w[4] = {q.x, q.y, q.z, q.w};
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            f[2 * k] = (float)(short)(w[k] & 0xFFFF) * (1.0f / 32768.0f);
            f[2 * k + 1] = (float)(short)(w[k] >> 16) * (1.0f / 32768.0f);
        }
        reinterpret_cast<float4 *>(y)[2 * v] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4 *>(y)[2 * v + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int64_t i = n8 * 8 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_samples; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = (float)x[i] * (1.0f / 32768.0f);
}

// ---------------------------------------------------------------------------------------------------------------
// One STFT column by one warp: iterative radix-2 DIT over M = N/2 complex points held in shared memory (zr, zi).
// `load(m)` returns sample m of the frame; the L band magnitudes (scaled) go to dst[0..L).
template <class LoadSample>
__device__ __forceinline__ void stft_band_column(const DevNet &net, LoadSample load, float *zr, float *zi, int lane, int bits,
                                                 float *__restrict__ dst, int scaling) {
    const int N = net.fft_len, M = N / 2, W = net.win_len, L = net.band;
    if (M == 1) {
        if (lane == 0) {
            float a = load(0) * net.window[0], b = W > 1 ? load(1) * net.window[1] : 0.0f;
            dst[0] = scale_value(fabsf(2.0f * (a + b)) / 2.0f, scaling);
        }
        return;
    }
    for (int n0 = lane; n0 < M; n0 += 4 * kWarp) {  // four points per lane per batch: loads first, then the stores
        float vr[4], vi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int n = n0 + k * kWarp, m0 = 2 * n, m1 = 2 * n + 1;
            vr[k] = (n < M && m0 < W) ? load(m0) * net.window[m0] : 0.0f;
            vi[k] = (n < M && m1 < W) ? load(m1) * net.window[m1] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int n = n0 + k * kWarp;
            if (n < M) {
                const int r = (int)(__brev((unsigned)n) >> (32 - bits));
                zr[r] = vr[k];
                zi[r] = vi[k];
            }
        }
    }
    __syncwarp();
    for (int lh = 0; (2 << lh) <= M; ++lh) {  // len = 2 << lh (a power of two: index maths by shifts)
        const int half = 1 << lh, ls = bits - lh;  // twiddle step = N / len = 1 << ls
        for (int idx = lane; idx < (M >> 1); idx += kWarp) {
            const int j = idx & (half - 1), a = ((idx >> lh) << (lh + 1)) + j, b = a + half;
            const float2 w = net.twiddle[j << ls];
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
        dst[f] = scale_value(mag, scaling);
    }
    __syncwarp();
}

__device__ __forceinline__ int log2_ceil(int m) {
    int bits = 0;
    while ((1 << bits) < m) ++bits;
    return bits;
}

// One warp per STFT column.
__global__ void stft_band_generic_kernel(const DevNet *__restrict__ netp, const float *__restrict__ pcm, int64_t ch_stride,
                                         int64_t col0, int64_t n_cols, float *__restrict__ feat, int64_t feat_ch_pitch,
                                         int scaling_override) {
    extern __shared__ float smem[];
    const DevNet &net = *netp;
    const int N = net.fft_len, M = N / 2, L = net.band;
    const int warps = blockDim.x / kWarp, warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    float *zr = smem + (size_t)warp * N, *zi = zr + M;
    const int ch = blockIdx.y;
    const float *x_ch = pcm + (int64_t)ch * ch_stride;
    float *feat_ch = feat + (int64_t)ch * feat_ch_pitch;
    const int bits = log2_ceil(M);
    const int scaling = scaling_override >= 0 ? scaling_override : net.scaling;

    for (int64_t c = (int64_t)blockIdx.x * warps + warp; c < n_cols; c += (int64_t)gridDim.x * warps) {
        const float *fr = x_ch + (col0 + c) * net.hop + net.gap;
        stft_band_column(net, [&](int m) { return fr[m]; }, zr, zi, lane, bits, feat_ch + c * L, scaling);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// High-overlap STFT for the wide-hidden tensor path (kernels_wide.cu): a CTA stages the audio span of kWideStftCols consecutive
// columns in shared memory ONCE (with hop 4 and a 1024-sample window, 64 columns share 1276 samples instead of reading 65 536), each
// warp then runs the reference-order FFT (stft_band_column) on frames of that span. The band magnitudes (after the spectrogram
// scaling) leave the kernel in the layout the tensor kernel's A operand wants: planes of 4 bins, [plane][row][4] float32, as a raw
// part (the tensor core truncates it to tf32) and a lo part (v - tf32(v)); plus per-column statistics for the per-window normalisers.
__device__ __forceinline__ float tf32_trunc_g(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(256) stft_planes_kernel(const DevNet *__restrict__ netp, const float *__restrict__ pcm, int64_t ch_stride,
                                                          int64_t col0, int64_t n_cols, float *__restrict__ hi, float *__restrict__ lo,
                                                          float4 *__restrict__ stats, int n_planes, int64_t rows_alloc) {
    extern __shared__ float smem[];
    const DevNet &net = *netp;
    const int N = net.fft_len, M = N / 2, L = net.band, W = net.win_len, hop = net.hop;
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int pitch = n_planes * 4 + 1;                       // output tile row pitch (odd: conflict-free column reads)
    const int span = (kWideStftCols - 1) * hop + W;
    float *audio = smem;                                       // [span]
    float *tile = audio + ((span + 3) & ~3);                   // [kWideStftCols][pitch]
    float *fft = tile + kWideStftCols * pitch;                 // [8 warps][N]
    float *zr = fft + (size_t)warp * N, *zi = zr + M;
    const int ch = blockIdx.y;
    const int64_t c_first = (int64_t)blockIdx.x * kWideStftCols;
    const int cols = (int)min((int64_t)kWideStftCols, n_cols - c_first);
    const float *src = pcm + (int64_t)ch * ch_stride + (col0 + c_first) * hop + net.gap;
    const int need = (cols - 1) * hop + W;
    for (int i = threadIdx.x; i < need; i += blockDim.x) audio[i] = src[i];
    for (int i = threadIdx.x; i < kWideStftCols * pitch; i += blockDim.x) tile[i] = 0.0f;   // padding bins and missing columns read as 0
    __syncthreads();
    const int bits = log2_ceil(M);
    for (int c = warp; c < cols; c += blockDim.x / kWarp) {
        const float *fr = audio + c * hop;
        float *row = tile + c * pitch;
        stft_band_column(net, [&](int m) { return fr[m]; }, zr, zi, lane, bits, row, net.scaling);
        __syncwarp();
        float ss = 0.0f, mn = INFINITY, mx = -INFINITY;
        for (int f = lane; f < L; f += kWarp) {
            const float v = row[f];
            ss += v * v;
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, d);
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        if (lane == 0) stats[(int64_t)ch * rows_alloc + c_first + c] = make_float4(ss, mn, mx, 0.0f);
    }
    __syncthreads();
    float4 *hi4 = reinterpret_cast<float4 *>(hi) + (int64_t)ch * n_planes * rows_alloc + c_first;
    float4 *lo4 = reinterpret_cast<float4 *>(lo) + (int64_t)ch * n_planes * rows_alloc + c_first;
    for (int i = threadIdx.x; i < n_planes * kWideStftCols; i += blockDim.x) {
        const int pl = i / kWideStftCols, r = i % kWideStftCols;   // consecutive threads: consecutive rows of one plane (16 B each)
        if (r < cols) {
            const float *v = tile + r * pitch + pl * 4;
            const float4 raw = make_float4(v[0], v[1], v[2], v[3]);
            hi4[(int64_t)pl * rows_alloc + r] = raw;
            lo4[(int64_t)pl * rows_alloc + r] = make_float4(raw.x - tf32_trunc_g(raw.x), raw.y - tf32_trunc_g(raw.y), raw.z - tf32_trunc_g(raw.z),
                                                            raw.w - tf32_trunc_g(raw.w));
        }
    }
}

// x[i] = f(i, x[i]) for the lane's elements, eight at a time: all loads of a batch are issued before its first store (the
// compiler cannot reorder them itself, x may alias whatever f reads).
template <class F>
__device__ __forceinline__ void lane_map(float *x, int n, int lane, F f) {
    for (int i = lane; i < n; i += 8 * kWarp) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i + k * kWarp < n) v[k] = f(i + k * kWarp, x[i + k * kWarp]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (i + k * kWarp < n) x[i + k * kWarp] = v[k];
    }
}

// acc = (((0 + t(0)) + t(1)) + ...) + t(n-1): the additions stay in index order (the order vDSP's scalar definition gives and
// the oracle uses); the terms of the next eight are fetched and formed while the current eight are being added, so the
// chain costs one FADD per element instead of a load-to-use latency per element.
template <class Term>
__device__ __forceinline__ float serial_sum(int n, Term term) {
    float acc = 0.0f;
    int i = 0;
    if (n >= 8) {
        float cur[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cur[k] = term(k);
        for (; i + 16 <= n; i += 8) {
            float nxt[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) nxt[k] = term(i + 8 + k);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += cur[k];
#pragma unroll
            for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += cur[k];
        i += 8;
    }
    for (; i < n; ++i) acc += term(i);
    return acc;
}

// The live tick trades the index-order sums for lane-strided partial sums combined by a shuffle butterfly (every lane ends
// with the same value): same operations, different association - within the stated 1e-5 output tolerance, and ~5x less
// latency per evaluation. The batch kernels below keep the index order (kSerial = true).
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
template <class Term>
__device__ __forceinline__ float strided_sum(int n, int lane, Term term) {
    float acc = 0.0f;
    for (int i = lane; i < n; i += 8 * kWarp) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = i + k * kWarp < n ? term(i + k * kWarp) : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    }
    return warp_sum(acc);
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per evaluation: literal processing chain, lanes own output neurons (each sum is left-to-right in one lane).
template <bool kSerial>
__device__ void apply_input_processing(const DevProcessing &p, float *x, int n, int lane) {
    switch (p.function) {
        case SYLDET_PROC_MAPMINMAX:
            lane_map(x, n, lane, [&](int i, float xi) { const float t = (xi - p.xoff[i]) * p.gain[i]; return t + p.y; });
            break;
        case SYLDET_PROC_MAPSTD:
            lane_map(x, n, lane, [&](int i, float xi) { const float t = (xi - p.xoff[i]) * p.gain[i]; return (0 != p.y) ? t + p.y : t; });
            break;
        case SYLDET_PROC_L2NORMALIZE: {
            float d = 0.0f;
            if constexpr (kSerial) {
                if (lane == 0) {
                    const float ss = serial_sum(n, [&](int i) { return x[i] * x[i]; });
                    d = sqrtf(ss);
                }
                d = __shfl_sync(0xffffffffu, d, 0);
            } else {
                d = sqrtf(strided_sum(n, lane, [&](int i) { return x[i] * x[i]; }));
            }
            lane_map(x, n, lane, [&](int, float xi) { return xi / d; });
            break;
        }
        case SYLDET_PROC_NORMALIZE: {
            float mn = 0.0f, mx = 0.0f;
            if constexpr (kSerial) {
                if (lane == 0) {
                    mn = mx = x[0];
                    for (int i = 1; i < n; ++i) { if (x[i] < mn) mn = x[i]; if (x[i] > mx) mx = x[i]; }
                }
                mn = __shfl_sync(0xffffffffu, mn, 0);
                mx = __shfl_sync(0xffffffffu, mx, 0);
            } else {
                mn = mx = x[0];
                for (int i = lane; i < n; i += kWarp) { const float v = x[i]; if (v < mn) mn = v; if (v > mx) mx = v; }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const float a = __shfl_xor_sync(0xffffffffu, mn, d), b = __shfl_xor_sync(0xffffffffu, mx, d);
                    if (a < mn) mn = a;
                    if (b > mx) mx = b;
                }
            }
            const float range = mx - mn;
            if (0 == range) {
                lane_map(x, n, lane, [&](int, float) { return -1.0f; });
            } else {
                const float slope = 2.0f / range, icpt = (0 - mn - mx) / range;
                lane_map(x, n, lane, [&](int, float xi) { const float t = xi * slope; return t + icpt; });
            }
            break;
        }
        case SYLDET_PROC_NORMALIZESTD: {
            float mean = 0.0f, sd = 0.0f;
            if constexpr (kSerial) {
                if (lane == 0) {
                    const float s = serial_sum(n, [&](int i) { return x[i]; });
                    mean = s / (float)n;
                    const float v = serial_sum(n, [&](int i) { const float d = x[i] - mean; return d * d; });
                    sd = sqrtf(v / (float)n);
                }
                mean = __shfl_sync(0xffffffffu, mean, 0);
                sd = __shfl_sync(0xffffffffu, sd, 0);
            } else {
                mean = strided_sum(n, lane, [&](int i) { return x[i]; }) / (float)n;
                sd = sqrtf(strided_sum(n, lane, [&](int i) { const float d = x[i] - mean; return d * d; }) / (float)n);
            }
            lane_map(x, n, lane, [&](int, float xi) { return (xi - mean) / sd; });
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

// One evaluation by one warp: `loadf(i)` returns input i (v[t*L+f], oldest column first). Returns the buffer holding the
// O reverse-mapped outputs (NeuralNet.apply, NeuralNet.swift:294-326).
template <bool kSerial, class LoadFeat>
__device__ __forceinline__ float *nn_evaluate(const DevNet &net, LoadFeat loadf, float *buf0, float *buf1, int lane, long long *dbg = nullptr) {
    const int I = net.inputs, O = net.outputs;
    float *cur = buf0, *nxt = buf1;
    for (int i = lane; i < I; i += 12 * kWarp) {  // gather in batches: every load of a batch is in flight before the first store
        float v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k)
            if (i + k * kWarp < I) v[k] = loadf(i + k * kWarp);
#pragma unroll
        for (int k = 0; k < 12; ++k)
            if (i + k * kWarp < I) cur[i + k * kWarp] = v[k];
    }
    __syncwarp();
    if (dbg) dbg[0] = clock64();
    for (int k = 0; k < net.n_ip; ++k) {
        apply_input_processing<kSerial>(net.ip[k], cur, I, lane);
        __syncwarp();
        if (dbg && k < 2) dbg[1 + k] = clock64();
    }
    for (int l = 0; l < net.n_layers; ++l) {
        const DevLayer &ly = net.layers[l];
        if constexpr (kSerial) {
            for (int o = lane; o < ly.outputs; o += kWarp) {
                const float *w = ly.w + (size_t)o * ly.inputs;
                const float acc = serial_sum(ly.inputs, [&](int i) { return w[i] * cur[i]; });
                nxt[o] = apply_transfer(ly.transfer, acc + ly.b[o]);
            }
        } else {  // four neurons at a time: the lanes split the inputs, lane k of the group finishes neuron o0 + k
            for (int o0 = 0; o0 < ly.outputs; o0 += 4) {
                float part[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                const float *w = ly.w + (size_t)o0 * ly.inputs;
                const int rows = min(4, ly.outputs - o0);
                for (int i = lane; i < ly.inputs; i += kWarp) {
                    const float xi = cur[i];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < rows) part[k] += w[(size_t)k * ly.inputs + i] * xi;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) part[k] = warp_sum(part[k]);
                if (lane < rows) {
                    const float acc = lane == 0 ? part[0] : lane == 1 ? part[1] : lane == 2 ? part[2] : part[3];
                    nxt[o0 + lane] = apply_transfer(ly.transfer, acc + ly.b[o0 + lane]);
                }
            }
        }
        __syncwarp();
        if (dbg && l < 2) dbg[3 + l] = clock64();
        float *t = cur; cur = nxt; nxt = t;
    }
    for (int o = lane; o < O; o += kWarp) {
        float v = cur[o];
        for (int k = 0; k < net.n_op; ++k) {  // reverse transforms, index order (NeuralNet.swift:316-323)
            const DevProcessing &p = net.op[k];
            float t = v + (0 - p.y);
            t = t / p.gain[o];
            v = t + p.xoff[o];
        }
        cur[o] = v;
    }
    __syncwarp();
    return cur;
}

__global__ void nn_generic_kernel(const DevNet *__restrict__ netp, const float *__restrict__ feat, int64_t n_cols,
                                  int64_t n_evals, int64_t eval0, int64_t evals_total, int detect_rule,
                                  float *__restrict__ all_out, EventSink sink) {
    extern __shared__ float smem[];
    const DevNet &net = *netp;
    const int warps = blockDim.x / kWarp, warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int O = net.outputs, L = net.band;
    float *buf0 = smem + (size_t)warp * 2 * net.max_width, *buf1 = buf0 + net.max_width;
    const int ch = blockIdx.y;
    const float *feat_ch = feat + (int64_t)ch * n_cols * L;

    for (int64_t j = (int64_t)blockIdx.x * warps + warp; j < n_evals; j += (int64_t)gridDim.x * warps) {
        const float *in = feat_ch + j * L;  // v[t*L+f]: contiguous in the column stream
        float *cur = nn_evaluate<true>(net, [&](int i) { return in[i]; }, buf0, buf1, lane);
        bool hit = false;
        for (int o = lane; o < O; o += kWarp) {
            const float v = cur[o];
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

// ---------------------------------------------------------------------------------------------------------------
// Live tick (Processor.swift:102-149 for every channel at once): ONE launch pulls the new samples of every channel
// straight out of pinned host memory into the channel's device sample ring, computes the STFT columns they complete
// into the channel's band-feature ring, evaluates the network for every completed feature window and writes the
// outputs straight into pinned host memory, then raises a host-visible sequence flag. Reference operations, unfused
// (this file is built with -fmad=false); the STFT keeps the reference order, the network's sums are lane-parallel. grid = (blocks per channel, channels); with all three phases in one launch
// there is one block per channel (block-level barriers order the phases); large ticks run one launch per phase.
__device__ __forceinline__ const float *relocate(const float *p, const unsigned char *from, const unsigned char *to) {
    return p ? reinterpret_cast<const float *>(to + (reinterpret_cast<const unsigned char *>(p) - from)) : nullptr;
}

__global__ void stream_tick_kernel(const DevNet *__restrict__ netp, StreamTick t) {
    extern __shared__ float smem[];
    __shared__ DevNet s_net;
    const int warps = blockDim.x / kWarp, warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int ch = blockIdx.y;
    const int tid0 = blockIdx.x * blockDim.x + threadIdx.x, tstride = gridDim.x * blockDim.x;
    const bool stamp = t.stamps != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;  // SYLDET_STREAM_TIMING
    long long ts[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (stamp) ts[0] = clock64();

    // (1) the longest-latency loads first: this channel's staged samples, straight out of pinned host memory
    const float *src = t.staged + (int64_t)ch * t.stage_pitch;
    float pre[2] = {0.0f, 0.0f};
    if ((t.phases & STREAM_PHASE_COPY) && !t.rs_on) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (tid0 + k * tstride < t.n_staged) pre[k] = src[tid0 + k * tstride];
    }
    // (2) while they fly: the configuration (DevNet record + window, twiddles, weights, processing vectors) -> shared
    // memory, so that the sums below never wait on L2. The blob (tens of KB) comes as ONE bulk copy (cp.async.bulk, completion on an
    // mbarrier) that runs while the samples arrive and are written to the ring; a per-thread copy loop of dependent L2 round trips
    // took 4 of the tick's 13 microseconds.
    __shared__ uint64_t blob_bar;
    unsigned char *sblob = reinterpret_cast<unsigned char *>(smem) + t.work_bytes;
    if (threadIdx.x == 0 && t.blob_bytes) {
        ptx::mbar_init(&blob_bar, 1);
        ptx::fence_mbar_init();
        ptx::mbar_expect_tx(&blob_bar, (uint32_t)t.blob_bytes);
        ptx::bulk_copy_g2s(sblob, t.blob, (uint32_t)t.blob_bytes, &blob_bar);
    }
    for (int i = threadIdx.x; i < (int)(sizeof(DevNet) / 4); i += blockDim.x)
        reinterpret_cast<int *>(&s_net)[i] = reinterpret_cast<const int *>(netp)[i];
    __syncthreads();
    if (t.blob_bytes && threadIdx.x == 0) {
        DevNet &n = s_net;
        for (int k = 0; k < n.n_ip; ++k) { n.ip[k].xoff = relocate(n.ip[k].xoff, t.blob, sblob); n.ip[k].gain = relocate(n.ip[k].gain, t.blob, sblob); }
        for (int k = 0; k < n.n_op; ++k) { n.op[k].xoff = relocate(n.op[k].xoff, t.blob, sblob); n.op[k].gain = relocate(n.op[k].gain, t.blob, sblob); }
        for (int l = 0; l < n.n_layers; ++l) { n.layers[l].w = relocate(n.layers[l].w, t.blob, sblob); n.layers[l].b = relocate(n.layers[l].b, t.blob, sblob); }
        n.window = relocate(n.window, t.blob, sblob);
        n.twiddle = reinterpret_cast<const float2 *>(relocate(reinterpret_cast<const float *>(n.twiddle), t.blob, sblob));
    }
    const DevNet &net = s_net;
    float *ring = t.ring + (int64_t)ch * (t.ring_mask + 1);
    if (stamp) ts[1] = clock64();

    if ((t.phases & STREAM_PHASE_COPY) && !t.rs_on) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (tid0 + k * tstride < t.n_staged) ring[(t.ring_pos + tid0 + k * tstride) & t.ring_mask] = pre[k];
        for (int i = tid0 + 2 * tstride; i < t.n_staged; i += tstride) ring[(t.ring_pos + i) & t.ring_mask] = src[i];
    } else if (t.phases & STREAM_PHASE_COPY) {
        // device-rate buffers -> ResamplerLinear -> sample ring: one resampleVector call per staged buffer (Processor.swift:116-121)
        for (int b = 0; b < t.n_marks; ++b) {
            const int lo = b ? t.marks[b - 1] : 0, n_in = t.marks[b] - lo, n_out = t.rs_n_out[b];
            const float off = t.rs_offset[b];
            const float last = b ? src[lo - 1] : t.rs_last_in[ch];
            for (int k = tid0; k < n_out; k += tstride)
                ring[(t.ring_pos + t.rs_out0[b] + k) & t.ring_mask] = resample_linear_point(src + lo, n_in, off, t.rs_step, last, off < 0.0f, k);
        }
        if (tid0 == 0 && t.n_staged > 0) t.rs_last_out[ch] = src[t.n_staged - 1];
    }
    if (t.blob_bytes) ptx::mbar_wait(&blob_bar, 0);   // the configuration has landed (the barrier was initialised before the sync above)
    __syncthreads();
    if (stamp) ts[2] = clock64();
    // Input level meter (Processor.swift:110-113, StatMax of the buffer's mean square): one warp per staged buffer, taken from the
    // back so that the warps the STFT columns do not need do it. Single-block launches read the samples back from the ring
    // (L2), the others from the staging area.
    if ((t.phases & STREAM_PHASE_COPY) && t.level_in && blockIdx.x == 0) {
        for (int b = 0; b < t.n_marks; ++b) {
            if (warps - 1 - (b % warps) != warp) continue;
            const int lo = b ? t.marks[b - 1] : 0, hi = t.marks[b];
            float part = 0.0f;
            for (int i = lo + lane; i < hi; i += kWarp) {
                const float v = (gridDim.x == 1 && !t.rs_on) ? ring[(t.ring_pos + i) & t.ring_mask] : src[i];   // the meter sees the device-rate samples
                part += v * v;
            }
            part = warp_sum(part);                                  // vDSP_svesq: summation order unspecified
            const double ms = (double)part / (double)(hi - lo);     // Double(sum) / Double(length)
            if (lane == 0 && hi > lo && ms == ms) atomicMax(t.level_in + ch, (unsigned long long)__double_as_longlong(ms));  // ms >= 0: bit order = value order
        }
    }
    const int L = net.band, O = net.outputs;
    float *band = t.band + (int64_t)ch * (t.band_mask + 1) * L;
    if (t.phases & STREAM_PHASE_COLUMNS) {
        const int N = net.fft_len, M = N / 2;
        float *zr = smem + (size_t)warp * N, *zi = zr + M;
        const int bits = log2_ceil(M);
        for (int64_t c = (int64_t)blockIdx.x * warps + warp; c < t.n_cols; c += (int64_t)gridDim.x * warps) {
            const int64_t first = (t.col0 + c) * net.hop + net.gap;
            stft_band_column(net, [&](int m) { return ring[(first + m) & t.ring_mask]; }, zr, zi, lane, bits,
                             band + ((t.col0 + c) & t.band_mask) * L, net.scaling);
        }
        __syncthreads();
    }
    if (stamp) ts[3] = clock64();
    if (t.phases & STREAM_PHASE_EVALS) {
        float *buf0 = smem + (size_t)warp * 2 * net.max_width, *buf1 = buf0 + net.max_width;
        for (int64_t j = (int64_t)blockIdx.x * warps + warp; j < t.n_evals; j += (int64_t)gridDim.x * warps) {
            const int64_t c0 = t.eval0 + j;  // oldest column of the window
            int tt = 0, f = lane;  // (column, bin) of input i = lane, lane + 32, ... without dividing
            float *cur = nn_evaluate<false>(net, [&](int) {
                while (f >= L) { f -= L; ++tt; }
                const float v = __ldcg(&band[((c0 + tt) & t.band_mask) * L + f]);  // written by earlier ticks or, moments ago, by this block: L2 is the coherent copy
                f += kWarp;
                return v;
            }, buf0, buf1, lane, stamp ? ts + 5 : nullptr);
            if (t.level_out && lane == 0) {  // output meter (Processor.swift:138, StatMax of Double(lastOutputs[0])); NaN never wins upstream either
                const float v0 = cur[0];
                const int bits = __float_as_int(v0);
                if (v0 == v0) atomicMax(t.level_out + ch, bits >= 0 ? bits : bits ^ 0x7fffffff);  // order-preserving map float -> int
            }
            if (t.packed) {  // one evaluation of <= 3 outputs: outputs and sequence number travel in one 16-byte store
                if (lane == 0)
                    t.packed[ch] = make_uint4(__float_as_uint(cur[0]), O > 1 ? __float_as_uint(cur[1]) : 0u, O > 2 ? __float_as_uint(cur[2]) : 0u, t.seq);
            } else {
                for (int o = lane; o < O; o += kWarp) t.out[((int64_t)ch * t.n_evals + j) * O + o] = cur[o];
            }
            __syncwarp();
        }
    }
    if (stamp) {
        ts[4] = clock64();
        for (int k = 0; k < 10; ++k) t.stamps[k] = ts[k];
    }
    if (t.flags && !t.packed) {  // publish: one flag per channel when the launch has one block per channel, else the last block raises flag 0
        __threadfence_system();
        __syncthreads();

        if (threadIdx.x == 0) {
            if (gridDim.x == 1) {
                *(volatile unsigned *)(t.flags + ch) = t.seq;
            } else if (atomicAdd(t.counter, 1u) == gridDim.x * gridDim.y - 1) {
                *t.counter = 0;
                __threadfence_system();
                for (unsigned c = 0; c < gridDim.y; ++c) *(volatile unsigned *)(t.flags + c) = t.seq;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Simulator trace (SyllableDetector/ViewControllerSimulator.swift:251-254, 308-344): one value per input sample,
// 0 until the first evaluation is due, then clamp(out0 / Float(thr0), 0, 1) of evaluation j held for one hop. A pure
// HBM stream: 4 B of network output per hop read, 2 or 4 B per sample written; a thread writes 8 consecutive samples.
__global__ void simulator_trace_kernel(const float *__restrict__ all_out, int64_t evals, int n_out, float thr0, int64_t first,
                                       int hop, int64_t n_samples, int format, void *__restrict__ trace, int64_t trace_stride) {
    const int ch = blockIdx.y;
    const float *out_ch = all_out + (int64_t)ch * evals * n_out;
    for (int64_t s0 = 8 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x); s0 < n_samples; s0 += 8 * (int64_t)gridDim.x * blockDim.x) {
        int64_t j = s0 >= first ? (s0 - first) / hop : -1;
        int64_t next = s0 >= first ? first + (j + 1) * hop : first;  // sample at which evaluation j + 1 takes over
        float v = 0.0f;
        bool fresh = true;
        for (int k = 0; k < 8 && s0 + k < n_samples; ++k) {
            const int64_t s = s0 + k;
            if (s >= next) { ++j; next += hop; fresh = true; }
            if (fresh) {
                fresh = false;
                v = 0.0f;
                if (j >= 0 && j < evals) {
                    v = out_ch[j * n_out] / thr0;
                    if (v > 1.0f) v = 1.0f;
                    else if (v < 0.0f) v = 0.0f;
                }
            }
            if (format == SYLDET_PCM_S16) {  // LPCM 16-bit writer settings (:203-211); NaN (silence through l2normalize) is written as 0
                const float q = v != v ? 0.0f : rintf(v * 32768.0f);
                ((int16_t *)trace)[(int64_t)ch * trace_stride + s] = (int16_t)(q > 32767.0f ? 32767.0f : q);
            } else {
                ((float *)trace)[(int64_t)ch * trace_stride + s] = v;
            }
        }
    }
}

}  // namespace

cudaError_t launch_ingest(const void *src, int format, int interleaved, int n_channels, int64_t n_samples, int64_t src_stride,
                          float *dst, int64_t dst_stride, cudaStream_t stream) {
    const int64_t total = (int64_t)n_channels * n_samples;
    if (total <= 0) return cudaSuccess;
    if (format == SYLDET_PCM_S16 && !interleaved && n_channels <= 65535) {
        const int64_t bx = std::min<int64_t>((n_samples / 8 + 255) / 256 + 1, std::max<int64_t>(1, 148 * 16 / n_channels));
        ingest_planar_s16_kernel<<<dim3((unsigned)bx, (unsigned)n_channels), 256, 0, stream>>>((const int16_t *)src, n_samples, src_stride, dst, dst_stride);
        return cudaGetLastError();
    }
    int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    ingest_kernel<<<blocks, 256, 0, stream>>>(src, format, interleaved, n_channels, n_samples, src_stride, dst, dst_stride);
    return cudaGetLastError();
}

cudaError_t launch_simulator_trace(const float *all_out, int n_channels, int64_t evals, int n_out, float thr0, int64_t first, int hop,
                                   int64_t n_samples, int format, void *trace, int64_t trace_stride, cudaStream_t stream) {
    if (n_samples <= 0 || n_channels <= 0) return cudaSuccess;
    const int64_t threads = (n_samples + 7) / 8;
    const int bx = (int)std::min<int64_t>((threads + 255) / 256, std::max<int64_t>(1, (148 * 16) / n_channels));
    dim3 grid((unsigned)bx, (unsigned)n_channels);
    simulator_trace_kernel<<<grid, 256, 0, stream>>>(all_out, evals, n_out, thr0, first, hop, n_samples, format, trace, trace_stride);
    return cudaGetLastError();
}

cudaError_t launch_stft_band_generic(const DevNet *d_net, int fft_len, const float *pcm, int64_t ch_stride, int n_channels,
                                     int64_t col0, int64_t n_cols, float *feat, int64_t feat_ch_pitch, int scaling_override,
                                     cudaStream_t stream) {
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
    stft_band_generic_kernel<<<grid, warps * 32, smem, stream>>>(d_net, pcm, ch_stride, col0, n_cols, feat, feat_ch_pitch, scaling_override);
    return cudaGetLastError();
}

size_t stft_planes_smem(int fft_len, int win_len, int hop, int n_planes) {
    const int span = (kWideStftCols - 1) * hop + win_len;
    return sizeof(float) * ((size_t)((span + 3) & ~3) + (size_t)kWideStftCols * (n_planes * 4 + 1) + (size_t)8 * fft_len);
}

cudaError_t launch_stft_planes(const DevNet *d_net, int fft_len, int win_len, int hop, const float *pcm, int64_t ch_stride, int n_channels,
                               int64_t col0, int64_t n_cols, float *hi, float *lo, float4 *stats, int n_planes, int64_t rows_alloc,
                               cudaStream_t stream) {
    if (n_cols <= 0 || n_channels <= 0) return cudaSuccess;
    const size_t smem = stft_planes_smem(fft_len, win_len, hop, n_planes);
    cudaError_t e = cudaFuncSetAttribute(stft_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((n_cols + kWideStftCols - 1) / kWideStftCols), (unsigned)n_channels);
    stft_planes_kernel<<<grid, 256, smem, stream>>>(d_net, pcm, ch_stride, col0, n_cols, hi, lo, stats, n_planes, rows_alloc);
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

namespace syldet {
size_t stream_tick_smem(int fft_len, int max_width, int *warps_out) {
    int warps = 4;
    auto need = [&](int w) { return (size_t)w * std::max<size_t>((size_t)fft_len, 2 * (size_t)max_width) * sizeof(float); };
    while (warps > 1 && need(warps) > 96 * 1024) warps >>= 1;
    if (warps_out) *warps_out = warps;
    return (need(warps) + 15) & ~(size_t)15;
}

cudaError_t launch_stream_tick(const DevNet *d_net, int fft_len, int max_width, int n_channels, int blocks_per_channel,
                               StreamTick t, cudaStream_t stream) {
    int warps = 4;
    const size_t work = stream_tick_smem(fft_len, max_width, &warps);
    t.work_bytes = (int)work;
    if (work + (size_t)t.blob_bytes > kStreamTickMaxSmem) t.blob_bytes = 0;  // too large to stage: read it through L2
    const size_t smem = work + (size_t)t.blob_bytes;
    if (smem > 48 * 1024) {  // opt-in size; per device, so not cached here (small configurations never get here)
        cudaError_t e = cudaFuncSetAttribute(stream_tick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamTickMaxSmem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((unsigned)blocks_per_channel, (unsigned)n_channels);
    stream_tick_kernel<<<grid, warps * 32, smem, stream>>>(d_net, t);
    return cudaGetLastError();
}
}  // namespace syldet
