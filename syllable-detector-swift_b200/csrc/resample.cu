// Batched sample-rate conversion on the device (north_star (1)):
//   linear     ResamplerLinear.resampleVector (Common/Resampler.swift:35-70), one call per channel over the whole buffer, bit-faithful
//              (float32 index ramp and all); the live path runs the same arithmetic inside stream_tick_kernel (resample.cuh).
//   polyphase  the CORRECT converter the reference leaves to AVFoundation for files ("audioSettings" asks the asset reader for
//              config.samplingRate, Common/SyllableDetector.swift:19-23) and regrets not having live ("Terrible quality",
//              Common/Resampler.swift:17-19): rational up / down by a Kaiser-windowed sinc FIR, the algorithm of
//              scipy.signal.resample_poly (firwin(20 max(up, down) + 1, 1 / max(up, down), ("kaiser", 5.0)) x up, zero-phase alignment,
//              zero padding at the ends), against which it is validated.
// Both kernels are HBM streams (4 B in + 4 B x ratio out per sample): a block stages the input span of its outputs in shared memory
// with 128-bit loads, every thread forms its outputs from shared memory and the block writes them with 128-bit stores.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "engine.hpp"
#include "resample.cuh"

namespace syldet {

namespace {

constexpr int kRsThreads = 256;
constexpr int kRsLinPerThread = 4;                       // outputs per thread of the linear kernel (one 128-bit store)
constexpr int kRsLinBlock = kRsThreads * kRsLinPerThread;

// Stages x[lo .. hi] (clamped to [0, n_in)) at s[0 ..]; s[i] = x[base + i] with base = lo rounded down to a multiple of 4 when the
// channel is 16-byte aligned (128-bit loads), else base = lo. Returns base. Out-of-range positions read as 0.
__device__ __forceinline__ int64_t stage_span(const float *__restrict__ x, int64_t n_in, int64_t lo, int64_t hi, float *s, bool aligned) {
    if (aligned) {
        const int64_t base = lo >= 0 ? (lo & ~(int64_t)3) : -((-lo + 3) & ~(int64_t)3);
        const int64_t n4 = (hi - base) / 4 + 1;
        for (int64_t v = threadIdx.x; v < n4; v += blockDim.x) {
            const int64_t i = base + 4 * v;
            float4 q;
            if (i >= 0 && i + 3 < n_in) q = __ldg(reinterpret_cast<const float4 *>(x + i));
            else q = make_float4(i >= 0 && i < n_in ? x[i] : 0.f, i + 1 >= 0 && i + 1 < n_in ? x[i + 1] : 0.f,
                                 i + 2 >= 0 && i + 2 < n_in ? x[i + 2] : 0.f, i + 3 >= 0 && i + 3 < n_in ? x[i + 3] : 0.f);
            reinterpret_cast<float4 *>(s)[v] = q;
        }
        return base;
    }
    for (int64_t i = lo + threadIdx.x; i <= hi; i += blockDim.x) s[i - lo] = i >= 0 && i < n_in ? x[i] : 0.f;
    return lo;
}

__global__ void __launch_bounds__(kRsThreads) resample_linear_batch_kernel(const float *__restrict__ in, int64_t in_stride, int64_t n_in, float step,
                                                                            float *__restrict__ out, int64_t out_stride, int64_t n_out) {
    extern __shared__ __align__(16) float s_lin[];
    const float *x = in + (int64_t)blockIdx.y * in_stride;
    float *y = out + (int64_t)blockIdx.y * out_stride;
    const bool in_al = (reinterpret_cast<uintptr_t>(x) & 15) == 0, out_al = (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    for (int64_t k0 = (int64_t)blockIdx.x * kRsLinBlock; k0 < n_out; k0 += (int64_t)gridDim.x * kRsLinBlock) {
        const int64_t k_last = min(k0 + kRsLinBlock, n_out) - 1;
        // fresh state: offset = 0 (Resampler.swift:24), so no blend across buffers; the float32 ramp is monotone in k
        const int64_t lo = (int64_t)__fmul_rn((float)k0, step), hi = min((int64_t)__fmul_rn((float)k_last, step) + 1, n_in - 1);
        __syncthreads();
        const int64_t base = stage_span(x, n_in, lo, hi, s_lin, in_al);
        __syncthreads();
        const int64_t k = k0 + (int64_t)threadIdx.x * kRsLinPerThread;
        float v[kRsLinPerThread];
#pragma unroll
        for (int j = 0; j < kRsLinPerThread; ++j) {
            v[j] = 0.0f;
            if (k + j < n_out) v[j] = resample_linear_point(s_lin - base, n_in, 0.0f, step, 0.0f, false, k + j);
        }
        if (out_al && k + kRsLinPerThread <= n_out) *reinterpret_cast<float4 *>(y + k) = make_float4(v[0], v[1], v[2], v[3]);
        else
            for (int j = 0; j < kRsLinPerThread; ++j)
                if (k + j < n_out) y[k + j] = v[j];
    }
}

// y[m] = sum_n x[n] hp[(m + r) down - n up]: per output the taps q0 + j up (q0 = ((m + r) down) mod up) against x[n_max - j].
__global__ void __launch_bounds__(kRsThreads) resample_poly_kernel(const float *__restrict__ in, int64_t in_stride, int64_t n_in, const float *__restrict__ hp,
                                                                    int len_hp, int up, int down, int64_t pre_remove, float *__restrict__ out,
                                                                    int64_t out_stride, int64_t n_out, int span_floats) {
    extern __shared__ __align__(16) float s_poly[];
    float *s_h = s_poly, *s_x = s_poly + ((len_hp + 3) & ~3);
    const float *x = in + (int64_t)blockIdx.y * in_stride;
    float *y = out + (int64_t)blockIdx.y * out_stride;
    const bool in_al = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    for (int i = threadIdx.x; i < len_hp; i += blockDim.x) s_h[i] = __ldg(hp + i);
    const int taps = (len_hp + up - 1) / up;
    for (int64_t m0 = (int64_t)blockIdx.x * kRsThreads; m0 < n_out; m0 += (int64_t)gridDim.x * kRsThreads) {
        const int64_t m_last = min(m0 + kRsThreads, n_out) - 1;
        const int64_t lo = ((m0 + pre_remove) * down) / up - (taps - 1), hi = min(((m_last + pre_remove) * down) / up, n_in - 1);
        __syncthreads();
        const int64_t base = stage_span(x, n_in, lo, hi, s_x, in_al);
        __syncthreads();
        const int64_t m = m0 + threadIdx.x;
        if (m < n_out) {
            const int64_t t = (m + pre_remove) * down;
            const int64_t n_max = t / up;
            const int q0 = (int)(t - n_max * up);
            const float *sx = s_x + (n_max - base);
            float acc = 0.0f;
            int j = 0;
            for (int q = q0; q < len_hp; q += up, ++j)
                if (n_max - j >= lo) acc = fmaf(s_h[q], n_max - j <= hi ? sx[-j] : 0.0f, acc);   // beyond either end of the signal: zero padding
            y[m] = acc;
        }
    }
    (void)span_floats;
}

double bessel_i0(double x) {   // power series; x <= 5 here
    double sum = 1.0, term = 1.0;
    for (int k = 1; k < 60; ++k) {
        term *= (x / (2.0 * k)) * (x / (2.0 * k));
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}

struct PolyDesign {
    int up = 1, down = 1;
    int64_t pre_remove = 0;
    std::vector<float> hp;   // zero-padded filter (scipy's h after the n_pre_pad zeros), x up
};

// scipy.signal.resample_poly's default design: firwin(2 half_len + 1, 1 / max_rate, window=("kaiser", 5.0)), scaled to unit DC gain, x up.
PolyDesign design_poly(int64_t rate_in, int64_t rate_out) {
    PolyDesign d;
    const int64_t g = std::gcd(rate_in, rate_out);
    d.up = (int)(rate_out / g);
    d.down = (int)(rate_in / g);
    const int max_rate = std::max(d.up, d.down);
    const double fc = 1.0 / max_rate;
    const int half_len = 10 * max_rate, n = 2 * half_len + 1;
    std::vector<double> h(n);
    const double alpha = 0.5 * (n - 1), beta = 5.0, i0b = bessel_i0(beta);
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double m = i - alpha, arg = fc * m;
        const double sinc = arg == 0.0 ? 1.0 : std::sin(M_PI * arg) / (M_PI * arg);
        const double r = m / alpha;
        const double w = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
        h[i] = fc * sinc * w;
        sum += h[i];
    }
    const int n_pre_pad = d.down - half_len % d.down;
    d.pre_remove = (half_len + n_pre_pad) / d.down;
    d.hp.assign((size_t)n_pre_pad + n, 0.0f);
    for (int i = 0; i < n; ++i) d.hp[(size_t)n_pre_pad + i] = (float)(h[i] / sum * d.up);
    return d;
}

bool integral_rate(double r, int64_t *out) {
    const double rr = std::nearbyint(r);
    if (!(r > 0.0) || std::fabs(r - rr) > 1e-9 || rr > 1e9) return false;
    *out = (int64_t)rr;
    return true;
}

}  // namespace

int64_t resample_output_length(int mode, int64_t n_in, double rate_in, double rate_out) {
    if (n_in <= 0 || !(rate_in > 0.0) || !(rate_out > 0.0)) return 0;
    if (mode == SYLDET_RESAMPLE_LINEAR) {
        const float step = (float)(rate_in / rate_out);   // Resampler.swift:32
        return std::max<int64_t>(0, (int64_t)((float)n_in / step));   // numSamplesOut with offset = 0 (:40)
    }
    int64_t a, b;
    if (!integral_rate(rate_in, &a) || !integral_rate(rate_out, &b)) return -1;
    const int64_t g = std::gcd(a, b), up = b / g, down = a / g;
    return (n_in * up + down - 1) / down;
}

// d_in / d_out: device pointers, planar. Asynchronous on `stream`.
syldet_status resample_device(int mode, const float *d_in, int n_channels, int64_t n_in, int64_t in_stride, double rate_in, double rate_out,
                              float *d_out, int64_t out_stride, int64_t *n_out, DeviceBuffer &filter, cudaStream_t stream) {
    if (!d_in || !d_out || n_channels <= 0 || n_channels > 65535 || n_in < 0) return set_error(SYLDET_ERR_ARG, "bad resampler arguments");
    if (mode != SYLDET_RESAMPLE_LINEAR && mode != SYLDET_RESAMPLE_POLYPHASE) return set_error(SYLDET_ERR_ARG, "unknown resampling mode");
    if (!(rate_in > 0.0) || !(rate_out > 0.0)) return set_error(SYLDET_ERR_ARG, "sampling rates must be positive");
    const int64_t n = resample_output_length(mode, n_in, rate_in, rate_out);
    if (n < 0) return set_error(SYLDET_ERR_UNSUPPORTED, "the polyphase converter needs integral sampling rates");
    if (n_out) *n_out = n;
    if (n == 0) return SYLDET_OK;
    if (out_stride < n || (n_channels > 1 && in_stride < n_in)) return set_error(SYLDET_ERR_ARG, "channel stride shorter than the channel");
    if (mode == SYLDET_RESAMPLE_LINEAR) {
        const float step = (float)(rate_in / rate_out);
        const size_t smem = ((size_t)((double)kRsLinBlock * step) + 16) * sizeof(float);
        if (smem > 200 * 1024) return set_error(SYLDET_ERR_UNSUPPORTED, "rate ratio too large for the batched linear resampler");
        SYLDET_CUDA(cudaFuncSetAttribute(resample_linear_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t blocks = std::min<int64_t>((n + kRsLinBlock - 1) / kRsLinBlock, std::max<int64_t>(1, 148 * 16 / n_channels));
        resample_linear_batch_kernel<<<dim3((unsigned)blocks, (unsigned)n_channels), kRsThreads, smem, stream>>>(d_in, in_stride, n_in, step, d_out,
                                                                                                                  out_stride, n);
        SYLDET_CUDA(cudaGetLastError());
        return SYLDET_OK;
    }
    int64_t a = 0, b = 0;
    integral_rate(rate_in, &a);
    integral_rate(rate_out, &b);
    const PolyDesign d = design_poly(a, b);
    if (d.up == 1 && d.down == 1) {
        SYLDET_CUDA(cudaMemcpy2DAsync(d_out, out_stride * 4, d_in, in_stride * 4, (size_t)n_in * 4, n_channels, cudaMemcpyDeviceToDevice, stream));
        return SYLDET_OK;
    }
    syldet_status st = filter.reserve(d.hp.size() * sizeof(float));
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaMemcpyAsync(filter.get(), d.hp.data(), d.hp.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    SYLDET_CUDA(cudaStreamSynchronize(stream));   // d.hp is a local
    const int len_hp = (int)d.hp.size(), taps = (len_hp + d.up - 1) / d.up;
    const int span = (int)((int64_t)kRsThreads * d.down / d.up) + taps + 16;
    const size_t smem = ((size_t)((len_hp + 3) & ~3) + span) * sizeof(float);
    if (smem > 200 * 1024) return set_error(SYLDET_ERR_UNSUPPORTED, "rate ratio too large for the polyphase resampler");
    SYLDET_CUDA(cudaFuncSetAttribute(resample_poly_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = std::min<int64_t>((n + kRsThreads - 1) / kRsThreads, std::max<int64_t>(1, 148 * 16 / n_channels));
    resample_poly_kernel<<<dim3((unsigned)blocks, (unsigned)n_channels), kRsThreads, smem, stream>>>(d_in, in_stride, n_in, filter.as<float>(), len_hp,
                                                                                                      d.up, d.down, d.pre_remove, d_out, out_stride, n, span);
    SYLDET_CUDA(cudaGetLastError());
    return SYLDET_OK;
}

syldet_status resample_host(int mode, const float *in, int n_channels, int64_t n_in, int64_t in_stride, double rate_in, double rate_out, float *out,
                            int64_t out_stride, int64_t *n_out, int device) {
    if (!in || !out) return set_error(SYLDET_ERR_ARG, "null argument");
    syldet_status st = use_device(device);
    if (st != SYLDET_OK) return st;
    const int64_t n = resample_output_length(mode, n_in, rate_in, rate_out);
    if (n < 0) return set_error(SYLDET_ERR_UNSUPPORTED, "the polyphase converter needs integral sampling rates");
    if (n_out) *n_out = n;
    if (n <= 0 || n_channels <= 0) return n_channels > 0 ? SYLDET_OK : set_error(SYLDET_ERR_ARG, "bad channel count");
    if (out_stride < n) return set_error(SYLDET_ERR_ARG, "output stride shorter than the output");
    const int64_t ip = (n_in + 3) & ~(int64_t)3, op = (n + 3) & ~(int64_t)3;
    DeviceBuffer d_in, d_out, filter;
    st = d_in.reserve((size_t)n_channels * ip * 4);
    if (st != SYLDET_OK) return st;
    st = d_out.reserve((size_t)n_channels * op * 4);
    if (st != SYLDET_OK) return st;
    cudaStream_t stream = nullptr;
    SYLDET_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    for (int ch = 0; ch < n_channels && e == cudaSuccess; ++ch)
        e = cudaMemcpyAsync(d_in.as<float>() + (size_t)ch * ip, in + (size_t)ch * (n_channels > 1 ? in_stride : n_in), (size_t)n_in * 4, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
        st = resample_device(mode, d_in.as<float>(), n_channels, n_in, ip, rate_in, rate_out, d_out.as<float>(), op, nullptr, filter, stream);
        for (int ch = 0; ch < n_channels && st == SYLDET_OK && e == cudaSuccess; ++ch)
            e = cudaMemcpyAsync(out + (size_t)ch * out_stride, d_out.as<float>() + (size_t)ch * op, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    }
    const cudaError_t e2 = cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    if (st != SYLDET_OK) return st;
    if (e != cudaSuccess) return cuda_fail(e, "resampler copies");
    if (e2 != cudaSuccess) return cuda_fail(e2, "resampler");
    return SYLDET_OK;
}

}  // namespace syldet
