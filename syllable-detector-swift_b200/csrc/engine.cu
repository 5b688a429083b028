#include "engine.hpp"

#include <cuda.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace syldet {

syldet_status cuda_fail(cudaError_t e, const char *what) {
    return set_error(SYLDET_ERR_CUDA, std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + what);
}

int usable_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

syldet_status use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(SYLDET_ERR_CUDA, "no CUDA device is visible; libsyldet_cuda has no CPU fallback");
    }
    if (device < 0 || device >= n) return set_error(SYLDET_ERR_ARG, "device index out of range");
    int major = 0;
    SYLDET_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return set_error(SYLDET_ERR_CUDA, "device is not compute capability 10.x; the kernels are built for sm_100a only");
    SYLDET_CUDA(cudaSetDevice(device));
    return SYLDET_OK;
}

DeviceBuffer::~DeviceBuffer() {
    if (ptr_) cudaFree(ptr_);
}

syldet_status DeviceBuffer::reserve(size_t bytes) {
    if (bytes <= size_) return SYLDET_OK;
    if (ptr_) { cudaFree(ptr_); ptr_ = nullptr; size_ = 0; }
    cudaError_t e = cudaMalloc(&ptr_, bytes);
    if (e != cudaSuccess) {
        ptr_ = nullptr;
        cudaGetLastError();
        return set_error(SYLDET_ERR_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    size_ = bytes;
    return SYLDET_OK;
}

// -----------------------------------------------------------------------------------------------------------------
FusedPlan plan_fused(const Config &c) {
    FusedPlan plan;
    auto no = [&](const std::string &why) { plan.ok = false; plan.why = why; return plan; };
    const int N = c.fourier_length;
    if (!fused_supports_fft(N)) return no("fourierLength not in {64,128,256,512}");
    if (c.band > kFusedMaxBand) return no("band wider than 128 bins");
    if ((int)c.layers.size() > kFusedMaxLayers) return no("more than 4 layers");
    if (c.outputs > kFusedMaxOut) return no("more than 8 outputs");
    if ((int)c.output_processing.size() > kMaxProcessing) return no("too many output processing functions");
    for (size_t l = 0; l < c.layers.size(); ++l)
        if (c.layers[l].outputs > kFusedMaxHidden) return no("a layer is wider than 8 units");
    const int H = c.layers[0].outputs, I = c.inputs;
    const int hp = H <= 4 ? 4 : 8;
    if ((long)I * hp > kFusedMaxW0) return no("layer 0 does not fit the kernel parameter space");

    // input chain: [one per-window statistic]? followed by per-position affine maps
    FusedParams &p = plan.params;
    std::memset(&p, 0, sizeof p);
    p.window_stat = FUSED_STAT_NONE;
    std::vector<double> A(I, 1.0), C(I, 0.0);  // y_i = A_i * (alpha x_i + beta) + C_i
    for (size_t k = 0; k < c.input_processing.size(); ++k) {
        const Processing &pr = c.input_processing[k];
        if (pr.function == SYLDET_PROC_MAPMINMAX || pr.function == SYLDET_PROC_MAPSTD) {
            for (int i = 0; i < I; ++i) {
                A[i] = A[i] * (double)pr.gains[i];
                C[i] = (C[i] - (double)pr.x_offsets[i]) * (double)pr.gains[i] + (double)pr.y;
            }
        } else {
            if (k != 0) return no("a per-window normaliser after another processing function");
            p.window_stat = pr.function == SYLDET_PROC_L2NORMALIZE ? FUSED_STAT_L2
                          : pr.function == SYLDET_PROC_NORMALIZE ? FUSED_STAT_MINMAX : FUSED_STAT_STD;
        }
    }
    const Layer &l0 = c.layers[0];
    for (int h = 0; h < H; ++h) {
        double v = 0.0, b = (double)l0.biases[h];
        for (int i = 0; i < I; ++i) {
            const double w = (double)l0.weights[(size_t)h * I + i];
            p.w0[(size_t)i * hp + h] = (float)(w * A[i]);
            v += w * A[i];
            b += w * C[i];
        }
        p.v[h] = (float)v;
        p.bprime[h] = (float)b;
    }
    p.n_layers = (int)c.layers.size();
    for (int l = 0; l < p.n_layers; ++l) p.tf[l] = c.layers[l].transfer;
    for (int l = 0; l < p.n_layers; ++l) p.width[l] = c.layers[l].outputs;
    for (int l = 1; l < p.n_layers; ++l) {
        const Layer &ly = c.layers[l];
        for (int o = 0; o < ly.outputs; ++o) {
            for (int i = 0; i < ly.inputs; ++i)
                p.rest_w[((l - 1) * kFusedMaxHidden + o) * kFusedMaxHidden + i] = ly.weights[(size_t)o * ly.inputs + i];
            p.rest_b[(l - 1) * kFusedMaxHidden + o] = ly.biases[o];
        }
    }
    p.n_out = c.outputs;
    p.n_op = (int)c.output_processing.size();
    for (int k = 0; k < p.n_op; ++k) {
        const Processing &pr = c.output_processing[k];
        p.op_y[k] = pr.y;
        for (int o = 0; o < c.outputs; ++o) {
            p.op_gain[k * kFusedMaxOut + o] = pr.gains[o];
            p.op_xoff[k * kFusedMaxOut + o] = pr.x_offsets[o];
        }
    }
    for (int o = 0; o < c.outputs; ++o) {
        p.thr[o] = c.thresholds[o];
        float t = (float)c.thresholds[o];
        if ((double)t < c.thresholds[o]) t = std::nextafterf(t, INFINITY);  // NaN thresholds stay NaN: never detected
        p.thr_f[o] = t;
    }

    p.win_len = c.window_length;
    p.gap = c.gap;
    p.hop = c.hop;
    p.k0 = c.k0;
    p.band = c.band;
    p.time_range = c.time_range;
    p.scaling = c.scaling;
    p.band_pitch = c.band | 1;
    const int rc = fused_round_cols(N);
    p.nn_tile = std::max(2 * rc, kFusedThreads - rc);
    p.ring_cols = p.nn_tile + 2 * rc + c.time_range;
    const long span = (long)(rc - 1) * c.hop + c.window_length + 3;
    if (span > (1L << 16)) return no("hop too large for shared-memory staging");
    p.abuf_floats = (int)((span + 3) & ~3L) + 4;
    plan.launch.fft_len = N;
    plan.launch.hp = hp;
    plan.launch.smem = fused_smem_bytes(N, p);
    if (plan.launch.smem > 200 * 1024) return no("shared-memory working set too large");
    plan.ok = true;
    return plan;
}

// -----------------------------------------------------------------------------------------------------------------
namespace {
float tf32_round(double v) {  // round to nearest, ties away (cvt.rna.tf32.f32): 10 explicit mantissa bits
    float f = (float)v;
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return f;
    u = (u + 0x1000u) & 0xFFFFE000u;
    std::memcpy(&f, &u, 4);
    return f;
}
uint16_t f32_to_f16_rn(float f) {  // round to nearest even, subnormals kept (cvt.rn.f16.f32)
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    x &= 0x7FFFFFFFu;
    if (x >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);   // >= 65536 (not reached: |values| <= 1)
    if (x < 0x38800000u) {                                      // below the smallest normal 2^-14: count units of 2^-24
        float a;
        std::memcpy(&a, &x, 4);
        return (uint16_t)(sign | (uint32_t)std::lrint((double)a * 16777216.0));
    }
    const uint32_t mant = x & 0x7FFFFFu, exp = (x >> 23) - 127 + 15;
    uint32_t half = (exp << 10) | (mant >> 13);
    const uint32_t rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) ++half;   // a carry moves into the exponent
    return (uint16_t)(sign | half);
}
float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, mant = h & 0x3FFu;
    float f;
    if (exp == 0) {
        f = (float)mant * (1.0f / 16777216.0f);   // subnormal: mant * 2^-24
        return sign ? -f : f;
    }
    const uint32_t u = sign | ((exp - 15 + 127) << 23) | (mant << 13);   // inf / NaN do not occur (|A| <= 1.08)
    std::memcpy(&f, &u, 4);
    return f;
}
}  // namespace

TcPlan plan_tc(const Config &c, const FusedPlan &fused) {
    TcPlan plan;
    auto no = [&](const std::string &why) { plan.ok = false; plan.why = why; return plan; };
    if (!fused.ok) return no("needs the fused epilogue: " + fused.why);
    const int kpad = tc_k_pad(), hop = c.hop, W = c.window_length, N = c.fourier_length;
    if (c.gap != 0) return no("gap configurations");
    if (hop % 4 != 0) return no("hop not a multiple of 4 samples (TMA row pitch)");
    if (hop < kpad - 8 || hop > kpad) return no("hop outside [128, 136]");
    if (W <= hop || W > 2 * hop) return no("window does not span exactly two hop rows");
    if (c.band > 32) return no("band wider than 32 bins");
    if (fused.params.window_stat == FUSED_STAT_STD) return no("normalizestd needs a second pass over the window");
    plan.params = fused.params;
    plan.hp = fused.launch.hp;
    plan.n0 = ((c.time_range * plan.hp + 15) / 16) * 16;
    if (!tc_layout_fits(c.time_range, plan.n0)) return no("time range x hidden width too large for the layer-0 product buffers");
    plan.smem = tc_smem_bytes(plan.params, plan.hp);
    if (plan.smem > 227 * 1024) return no("shared-memory working set too large");
    // layer-0 weights as the B operand of the second contraction: Wcat[(t, h)][f] = W'[t*L + f][h], hi / lo
    plan.wcat_hi.assign((size_t)plan.n0 * 32, 0.0f);
    plan.wcat_lo.assign((size_t)plan.n0 * 32, 0.0f);
    for (int t = 0; t < c.time_range; ++t)
        for (int h = 0; h < plan.hp; ++h)
            for (int f = 0; f < c.band; ++f) {
                const double v = (double)fused.params.w0[(size_t)(t * c.band + f) * plan.hp + h];
                const size_t i = (size_t)(t * plan.hp + h) * 32 + f;
                plan.wcat_hi[i] = tf32_round(v);
                plan.wcat_lo[i] = tf32_round(v - (double)plan.wcat_hi[i]);
            }
    // rows: [0,32) Re B1 | [32,64) Im B1 | [64,96) Re B2 | [96,128) Im B2 ; B1 = samples [0,hop), B2 = samples [hop,W)
    plan.dft_hi.assign((size_t)128 * kpad, 0.0f);
    plan.dft_lo.assign((size_t)128 * kpad, 0.0f);
    // fp16 A operand of the kF16 variant (kernels_tc.cu): per row 416 fp16 = 208 words, K order
    //   [a1(n < 128) | a1 2^-11 (n < 128) | a1(n >= 128) (8) | a1 2^-11 (n >= 128) (8) | a2(n < 128) | a2(n >= 128) (8) | 0 (8)]
    // with a1 = fp16(A), a2 = fp16(A - a1); two k per 32-bit word (low half first)
    const int a16_halves = tc_a16_cols() * 2;
    std::vector<uint16_t> a16((size_t)128 * a16_halves, 0);
    auto put16 = [&](size_t row, int n, double full) {
        const uint16_t h1 = f32_to_f16_rn((float)full);
        const float a1 = f16_to_f32(h1);
        const bool main = n < 128;
        const int k = n & 127;
        uint16_t *dst = a16.data() + row * a16_halves;
        dst[main ? k : 256 + k] = h1;
        dst[main ? 128 + k : 264 + k] = f32_to_f16_rn(a1 * (1.0f / 2048.0f));
        dst[main ? 272 + k : 400 + k] = f32_to_f16_rn((float)(full - (double)a1));
    };
    for (int b = 0; b < c.band; ++b) {
        const int k = c.k0 + b;
        for (int half = 0; half < 2; ++half)
            for (int n = 0; n < hop; ++n) {
                const int m = half * hop + n;  // sample index inside the frame
                if (m >= W) break;
                const double win = (double)(float)(0.54 - 0.46 * std::cos(2.0 * M_PI * (double)m / (double)W));
                const double ang = -2.0 * M_PI * (double)(((long)k * m) % N) / (double)N;
                const double re = win * std::cos(ang), im = win * std::sin(ang);
                const size_t ire = (size_t)(half * 64 + b) * kpad + n, iim = (size_t)(half * 64 + 32 + b) * kpad + n;
                plan.dft_hi[ire] = tf32_round(re);
                plan.dft_lo[ire] = tf32_round(re - (double)plan.dft_hi[ire]);
                plan.dft_hi[iim] = tf32_round(im);
                plan.dft_lo[iim] = tf32_round(im - (double)plan.dft_hi[iim]);
                put16((size_t)(half * 64 + b), n, re);
                put16((size_t)(half * 64 + 32 + b), n, im);
            }
    }
    plan.dft16.assign((size_t)128 * tc_a16_cols(), 0u);
    for (size_t i = 0; i < plan.dft16.size(); ++i) plan.dft16[i] = (uint32_t)a16[2 * i] | ((uint32_t)a16[2 * i + 1] << 16);
    plan.ok = true;
    return plan;
}

// -----------------------------------------------------------------------------------------------------------------
WidePlan plan_wide(const Config &c) {
    WidePlan plan;
    auto no = [&](const std::string &why) { plan.ok = false; plan.why = why; return plan; };
    if (c.gap != 0 || c.hop != 4) return no("hop is not 4 samples (the sliding A operand needs 16-byte rows)");
    if (c.layers.size() != 2) return no("not a two-layer network");
    const int H = c.layers[0].outputs, I = c.inputs, O = c.outputs, T = c.time_range, L = c.band;
    if (H < 16 || H > 1024) return no("hidden width outside [16, 1024]");
    if (O > kFusedMaxOut) return no("more than 8 outputs");
    if (T > wide_max_time_range()) return no("time range above 17");
    if ((int)c.output_processing.size() > kMaxProcessing) return no("too many output processing functions");
    if (c.fourier_length > 4096) return no("fourierLength above 4096 (shared-memory staging of stft_planes_kernel)");
    WideParams &p = plan.params;
    std::memset(&p, 0, sizeof p);
    // input chain: [one per-window statistic]? followed by per-position affine maps: y_i = A_i (alpha x_i + beta) + C_i (as plan_fused)
    p.window_stat = FUSED_STAT_NONE;
    std::vector<double> A(I, 1.0), C(I, 0.0);
    for (size_t k = 0; k < c.input_processing.size(); ++k) {
        const Processing &pr = c.input_processing[k];
        if (pr.function == SYLDET_PROC_MAPMINMAX || pr.function == SYLDET_PROC_MAPSTD) {
            for (int i = 0; i < I; ++i) {
                A[i] = A[i] * (double)pr.gains[i];
                C[i] = (C[i] - (double)pr.x_offsets[i]) * (double)pr.gains[i] + (double)pr.y;
            }
        } else {
            if (k != 0) return no("a per-window normaliser after another processing function");
            if (pr.function == SYLDET_PROC_NORMALIZESTD) return no("normalizestd needs a second pass over the window");
            p.window_stat = pr.function == SYLDET_PROC_L2NORMALIZE ? FUSED_STAT_L2 : FUSED_STAT_MINMAX;
        }
    }
    p.time_range = T;
    p.band = L;
    const int planes = (L + 3) / 4;
    p.n_planes = ((planes + kWidePL - 1) / kWidePL) * kWidePL;
    p.hidden = H;
    p.h_pad = ((H + 255) / 256) * 256;
    p.n_out = O;
    p.n_op = (int)c.output_processing.size();
    p.tf0 = c.layers[0].transfer;
    p.tf1 = c.layers[1].transfer;
    for (int o = 0; o < O; ++o) p.b1[o] = c.layers[1].biases[o];
    for (int k = 0; k < p.n_op; ++k) {
        const Processing &pr = c.output_processing[k];
        p.op_y[k] = pr.y;
        for (int o = 0; o < O; ++o) {
            p.op_gain[k * kFusedMaxOut + o] = pr.gains[o];
            p.op_xoff[k * kFusedMaxOut + o] = pr.x_offsets[o];
        }
    }
    for (int o = 0; o < O; ++o) {
        float t = (float)c.thresholds[o];
        if ((double)t < c.thresholds[o]) t = std::nextafterf(t, INFINITY);
        p.thr_f[o] = t;
    }
    if (wide_smem_bytes(p.h_pad) > 227 * 1024) return no("shared-memory working set too large");
    // folded constants
    const Layer &l0 = c.layers[0], &l1 = c.layers[1];
    plan.v.assign(p.h_pad, 0.0f);
    plan.bprime.assign(p.h_pad, 0.0f);
    std::vector<double> wf((size_t)H * I);
    for (int h = 0; h < H; ++h) {
        double v = 0.0, b = (double)l0.biases[h];
        for (int i = 0; i < I; ++i) {
            const double w = (double)l0.weights[(size_t)h * I + i];
            wf[(size_t)h * I + i] = w * A[i];
            v += w * A[i];
            b += w * C[i];
        }
        plan.v[h] = (float)v;
        plan.bprime[h] = (float)b;
    }
    plan.w1.assign((size_t)O * p.h_pad, 0.0f);
    for (int o = 0; o < O; ++o)
        for (int h = 0; h < H; ++h) plan.w1[(size_t)o * p.h_pad + h] = l1.weights[(size_t)o * H + h];
    // weight blocks in streaming order: [pass nc][chunk c][t][hi | lo][plane q][n < 256][e < 4], input index i = t*L + f,
    // f = (c*kWidePL + q)*4 + e (zero for f >= L and for padded hidden units)
    const int n_nc = p.h_pad / 256, n_chunks = p.n_planes / kWidePL;
    const size_t blk = wide_weight_block_bytes() / sizeof(float), half = blk / 2;
    plan.weights.assign((size_t)n_nc * n_chunks * T * blk, 0.0f);
    for (int nc = 0; nc < n_nc; ++nc)
        for (int cch = 0; cch < n_chunks; ++cch)
            for (int t = 0; t < T; ++t) {
                float *dst = plan.weights.data() + ((size_t)(nc * n_chunks + cch) * T + t) * blk;
                for (int q = 0; q < kWidePL; ++q)
                    for (int n = 0; n < 256; ++n) {
                        const int h = nc * 256 + n;
                        if (h >= H) continue;
                        for (int e = 0; e < 4; ++e) {
                            const int f = (cch * kWidePL + q) * 4 + e;
                            if (f >= L) continue;
                            const double w = wf[(size_t)h * I + (size_t)t * L + f];
                            const float hi = tf32_round(w);
                            const size_t idx = ((size_t)q * 256 + n) * 4 + e;
                            dst[idx] = hi;
                            dst[half + idx] = tf32_round(w - (double)hi);
                        }
                    }
            }
    plan.ok = true;
    return plan;
}

// -----------------------------------------------------------------------------------------------------------------
syldet_status DeviceModel::init(const Config &cfg, int device) {
    cfg_ = cfg;
    if (!cfg_.valid) {
        syldet_status st = validate_config(cfg_);
        if (st != SYLDET_OK) return st;
    }
    if ((int)cfg_.input_processing.size() > kMaxProcessing || (int)cfg_.output_processing.size() > kMaxProcessing)
        return set_error(SYLDET_ERR_UNSUPPORTED, "more than 8 processing functions in a chain");
    if ((int)cfg_.layers.size() > kMaxLayers) return set_error(SYLDET_ERR_UNSUPPORTED, "more than 8 layers");
    if (cfg_.fourier_length > 16384) return set_error(SYLDET_ERR_UNSUPPORTED, "fourierLength above 16384");
    max_width_ = cfg_.inputs;
    for (const Layer &l : cfg_.layers) max_width_ = std::max(max_width_, l.outputs);
    if (max_width_ > 12288) return set_error(SYLDET_ERR_UNSUPPORTED, "a layer wider than 12288 units");
    syldet_status st = use_device(device);
    if (st != SYLDET_OK) return st;
    device_ = device;
    SYLDET_CUDA(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device));

    // one blob: [thresholds (double)] [window] [twiddle] [per-layer w,b] [per-processing xoff,gain]
    std::vector<unsigned char> blob;
    auto put = [&](const void *src, size_t bytes) {
        size_t off = (blob.size() + 15) & ~(size_t)15;
        blob.resize(off + bytes);
        std::memcpy(blob.data() + off, src, bytes);
        return off;
    };
    const int N = cfg_.fourier_length, W = cfg_.window_length;
    std::vector<float> window(W);
    for (int n = 0; n < W; ++n)  // vDSP_hamm_window, N-denominator (CSTFT.swift:24, forced at SyllableDetector.swift:43)
        window[n] = (float)(0.54 - 0.46 * std::cos(2.0 * M_PI * (double)n / (double)W));
    std::vector<float2> tw(std::max(1, N / 2));
    for (int k = 0; k < N / 2; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)N;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    const size_t off_thr = put(cfg_.thresholds.data(), cfg_.thresholds.size() * sizeof(double));
    const size_t off_win = put(window.data(), window.size() * sizeof(float));
    const size_t off_tw = put(tw.data(), tw.size() * sizeof(float2));
    std::vector<size_t> off_w, off_b;
    for (const Layer &l : cfg_.layers) {
        off_w.push_back(put(l.weights.data(), l.weights.size() * sizeof(float)));
        off_b.push_back(put(l.biases.data(), l.biases.size() * sizeof(float)));
    }
    auto put_proc = [&](const std::vector<Processing> &ps, std::vector<size_t> &xo, std::vector<size_t> &g) {
        for (const Processing &p : ps) {
            xo.push_back(p.x_offsets.empty() ? 0 : put(p.x_offsets.data(), p.x_offsets.size() * sizeof(float)));
            g.push_back(p.gains.empty() ? 0 : put(p.gains.data(), p.gains.size() * sizeof(float)));
        }
    };
    std::vector<size_t> ixo, ig, oxo, og;
    put_proc(cfg_.input_processing, ixo, ig);
    put_proc(cfg_.output_processing, oxo, og);

    blob.resize((blob.size() + 15) & ~(size_t)15);
    blob_bytes_ = blob.size();
    st = d_blob_.reserve(blob.size());
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaMemcpy(d_blob_.get(), blob.data(), blob.size(), cudaMemcpyHostToDevice));
    const unsigned char *base = d_blob_.as<unsigned char>();
    d_window_ = reinterpret_cast<const float *>(base + off_win);
    d_twiddle_ = reinterpret_cast<const float2 *>(base + off_tw);

    DevNet net{};
    net.fft_len = N;
    net.win_len = W;
    net.gap = cfg_.gap;
    net.hop = cfg_.hop;
    net.k0 = cfg_.k0;
    net.band = cfg_.band;
    net.time_range = cfg_.time_range;
    net.inputs = cfg_.inputs;
    net.outputs = cfg_.outputs;
    net.scaling = cfg_.scaling;
    net.n_ip = (int)cfg_.input_processing.size();
    net.n_op = (int)cfg_.output_processing.size();
    net.n_layers = (int)cfg_.layers.size();
    net.max_width = max_width_;
    auto fill_proc = [&](const std::vector<Processing> &ps, const std::vector<size_t> &xo, const std::vector<size_t> &g, DevProcessing *dst) {
        for (size_t i = 0; i < ps.size(); ++i) {
            dst[i].function = ps[i].function;
            dst[i].y = ps[i].y;
            dst[i].xoff = ps[i].x_offsets.empty() ? nullptr : reinterpret_cast<const float *>(base + xo[i]);
            dst[i].gain = ps[i].gains.empty() ? nullptr : reinterpret_cast<const float *>(base + g[i]);
        }
    };
    fill_proc(cfg_.input_processing, ixo, ig, net.ip);
    fill_proc(cfg_.output_processing, oxo, og, net.op);
    for (size_t i = 0; i < cfg_.layers.size(); ++i) {
        net.layers[i].inputs = cfg_.layers[i].inputs;
        net.layers[i].outputs = cfg_.layers[i].outputs;
        net.layers[i].transfer = cfg_.layers[i].transfer;
        net.layers[i].w = reinterpret_cast<const float *>(base + off_w[i]);
        net.layers[i].b = reinterpret_cast<const float *>(base + off_b[i]);
    }
    net.thresholds = reinterpret_cast<const double *>(base + off_thr);
    net.window = d_window_;
    net.twiddle = d_twiddle_;
    st = d_net_.reserve(sizeof(DevNet));
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaMemcpy(d_net_.get(), &net, sizeof net, cudaMemcpyHostToDevice));

    fused_ = plan_fused(cfg_);
    if (fused_.ok) {
        int blocks = 0;
        cudaError_t e = fused_max_blocks_per_sm(fused_.launch.fft_len, fused_.launch.hp, fused_.launch.smem, &blocks);
        if (e != cudaSuccess || blocks < 1) {
            cudaGetLastError();
            fused_.ok = false;
            fused_.why = std::string("fused kernel cannot be resident: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 blocks per SM");
        } else {
            fused_.blocks_per_sm = blocks;
            st = d_fused_params_.reserve(sizeof(FusedParams));
            if (st != SYLDET_OK) return st;
            SYLDET_CUDA(cudaMemcpy(d_fused_params_.get(), &fused_.params, sizeof(FusedParams), cudaMemcpyHostToDevice));
        }
    }
    wide_ = plan_wide(cfg_);
    if (wide_.ok) {
        const size_t nw = wide_.weights.size(), nh = wide_.v.size(), n1 = wide_.w1.size();
        st = d_wide_.reserve((nw + 2 * nh + n1) * sizeof(float));
        if (st != SYLDET_OK) return st;
        float *d = d_wide_.as<float>();
        SYLDET_CUDA(cudaMemcpy(d, wide_.weights.data(), nw * sizeof(float), cudaMemcpyHostToDevice));
        SYLDET_CUDA(cudaMemcpy(d + nw, wide_.v.data(), nh * sizeof(float), cudaMemcpyHostToDevice));
        SYLDET_CUDA(cudaMemcpy(d + nw + nh, wide_.bprime.data(), nh * sizeof(float), cudaMemcpyHostToDevice));
        SYLDET_CUDA(cudaMemcpy(d + nw + 2 * nh, wide_.w1.data(), n1 * sizeof(float), cudaMemcpyHostToDevice));
    }
    tc_ = plan_tc(cfg_, fused_);
    if (tc_.ok) {
        const size_t n = tc_.dft_hi.size(), nw = tc_.wcat_hi.size();
        st = d_dft_.reserve((2 * n + 2 * nw + tc_.dft16.size()) * sizeof(float));
        if (st != SYLDET_OK) return st;
        SYLDET_CUDA(cudaMemcpy(d_dft_.get(), tc_.dft_hi.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        SYLDET_CUDA(cudaMemcpy(d_dft_.as<float>() + n, tc_.dft_lo.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        SYLDET_CUDA(cudaMemcpy(d_dft_.as<float>() + 2 * n, tc_.wcat_hi.data(), nw * sizeof(float), cudaMemcpyHostToDevice));
        SYLDET_CUDA(cudaMemcpy(d_dft_.as<float>() + 2 * n + nw, tc_.wcat_lo.data(), nw * sizeof(float), cudaMemcpyHostToDevice));
        SYLDET_CUDA(cudaMemcpy(d_dft_.as<float>() + 2 * n + 2 * nw, tc_.dft16.data(), tc_.dft16.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    return SYLDET_OK;
}

// -----------------------------------------------------------------------------------------------------------------
void debounce_sorted(const Config &cfg, std::vector<syldet_event> &rows, std::vector<float> &outputs, int n_out,
                     int64_t debounce_frames) {
    // rows sorted by (channel, sample). emit iff debounceUntil < S; then debounceUntil = S + D  (TrackDetector.swift:80,99)
    (void)cfg;
    if (debounce_frames == 0) return;   // rows of a channel have strictly increasing sample numbers: every row is emitted
    size_t w = 0;
    int32_t cur_ch = INT32_MIN;
    int64_t until = -1;
    for (size_t r = 0; r < rows.size(); ++r) {
        if (rows[r].channel != cur_ch) { cur_ch = rows[r].channel; until = -1; }
        if (until < rows[r].sample) {
            until = rows[r].sample + debounce_frames;
            if (w != r) {
                rows[w] = rows[r];
                std::copy(outputs.begin() + r * n_out, outputs.begin() + (r + 1) * n_out, outputs.begin() + w * n_out);
            }
            ++w;
        }
    }
    rows.resize(w);
    outputs.resize(w * n_out);
}

syldet_status Batch::init(const Config &cfg, int device) {
    syldet_status st = model_.init(cfg, device);
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    int mp = 0;
    if (cudaDeviceGetAttribute(&mp, cudaDevAttrMaxPitch, device) == cudaSuccess && mp > 0) max_pitch_ = (size_t)mp;
    st = sink_count_.reserve(kSinkHeaderBytes);   // [0, 8): detection count, [8, 12): range flag of the fp16 correction pass
    return st;
}

Batch::~Batch() {
    if (own_stream_) {
        cudaSetDevice(model_.device());
        cudaStreamDestroy(own_stream_);
    }
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (d2h_stream_) cudaStreamDestroy(d2h_stream_);
    for (cudaEvent_t e : ev_copied_) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_done_) cudaEventDestroy(e);
    for (cudaEvent_t e : wide_ev_) cudaEventDestroy(e);
    if (h_counts_) cudaFreeHost(h_counts_);
    if (h_events_) cudaFreeHost(h_events_);
}

syldet_status Batch::set_kernel(int kernel) {
    if (kernel != SYLDET_KERNEL_AUTO && kernel != SYLDET_KERNEL_GENERIC && kernel != SYLDET_KERNEL_FUSED && kernel != SYLDET_KERNEL_TENSOR &&
        kernel != SYLDET_KERNEL_TENSOR_TF32 && kernel != SYLDET_KERNEL_WIDE)
        return set_error(SYLDET_ERR_ARG, "unknown kernel selector");
    if (kernel == SYLDET_KERNEL_WIDE && !model_.wide().ok)
        return set_error(SYLDET_ERR_UNSUPPORTED, "wide-hidden tensor kernel not available for this configuration: " + model_.wide().why);
    if ((kernel == SYLDET_KERNEL_TENSOR || kernel == SYLDET_KERNEL_TENSOR_TF32) && !model_.tc().ok)
        return set_error(SYLDET_ERR_UNSUPPORTED, "tensor-core kernel not available for this configuration: " + model_.tc().why);
    if (kernel == SYLDET_KERNEL_FUSED && !model_.fused().ok)
        return set_error(SYLDET_ERR_UNSUPPORTED, "fused kernel not available for this configuration: " + model_.fused().why);
    kernel_ = kernel;
    return SYLDET_OK;
}

int Batch::active_kernel() const {
    if (kernel_ == SYLDET_KERNEL_AUTO)
        return model_.tc().ok ? SYLDET_KERNEL_TENSOR : model_.fused().ok ? SYLDET_KERNEL_FUSED : model_.wide().ok ? SYLDET_KERNEL_WIDE : SYLDET_KERNEL_GENERIC;
    return kernel_;
}

syldet_status Batch::ensure_sink(unsigned long long capacity) {
    if (capacity <= sink_capacity_) return SYLDET_OK;
    syldet_status st = sink_events_.reserve(capacity * sizeof(DevEvent));
    if (st != SYLDET_OK) return st;
    st = sink_outputs_.reserve(capacity * sizeof(float) * model_.config().outputs);
    if (st != SYLDET_OK) return st;
    sink_capacity_ = capacity;
    return SYLDET_OK;
}

// Unit length of the tensor kernel in tiles. Units go to the CTAs round-robin, so the launch takes ceil(units / CTAs) units on the busiest
// CTA: pick the (even: tiles are contracted in pairs) length that minimises that, one tile charged per unit for the pipeline fill and
// the warm-up columns; a unit must be longer than four warm-ups. 1 h x 8 ch on 148 CTAs: 32 tiles per unit leaves 4 800 units = 32.4 per
// CTA, i.e. 33 on some (2.3 % over the ideal); 74 gives 2 072 = 14 each.
int tc_plan_unit_tiles(int64_t eval_count, int n_channels, int resident, int tile_frames, int warm) {
    const int64_t tf = tile_frames;
    int64_t best_tpu = 32;
    double best = 1e300;
    for (int64_t tpu = 4; tpu <= 96; tpu += 2) {
        if (tpu * tf <= 4 * (int64_t)warm) continue;
        const int64_t ch = tpu * tf - warm;
        const int64_t units = (int64_t)n_channels * ((std::max<int64_t>(eval_count, 1) + ch - 1) / ch);
        const double cost = (double)((units + resident - 1) / resident) * (double)(tpu + 1);
        if (cost < best) { best = cost; best_tpu = tpu; }
    }
    while (best_tpu * tf <= 4 * (int64_t)warm) best_tpu *= 2;   // very long windows: no candidate above qualified
    return (int)best_tpu;
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
}  // namespace

syldet_status Batch::launch_fused_range(const float *d_planar, int n_channels, int64_t ch_stride, const float *valid_begin,
                                        const float *valid_end, int64_t eval_begin, int64_t eval_count, int64_t evals_total,
                                        int detect_rule, float *d_all_outputs, EventSink sink, cudaStream_t stream) {
    const FusedPlan &fp = model_.fused();
    FusedWork w{};
    w.pcm = d_planar;
    w.pcm_begin = valid_begin;
    w.pcm_end = valid_end;
    w.ch_stride = ch_stride;
    w.n_channels = n_channels;
    w.evals_per_channel = eval_count;
    w.eval_offset = eval_begin;
    w.out_evals_per_channel = evals_total;
    const int resident = model_.sm_count() * fp.blocks_per_sm;
    // chunk: a multiple of nn_tile, long enough that the T-1 warm-up columns are noise, short enough for >= 8 waves
    const int64_t tile = fp.params.nn_tile;
    int64_t chunk = ((eval_count + tile - 1) / tile) * tile;
    const int64_t want_units = (int64_t)resident * 8;
    while (chunk > 4 * tile && n_channels * ((eval_count + chunk - 1) / chunk) < want_units) chunk = ((chunk / 2 + tile - 1) / tile) * tile;
    if (chunk > 64 * tile) chunk = 64 * tile;
    w.chunk_evals = chunk;
    w.chunks_per_channel = (int)((eval_count + chunk - 1) / chunk);
    w.detect_rule = detect_rule;
    w.all_out = d_all_outputs;
    w.sink = sink;
    w.window = model_.window();
    w.twiddle = model_.twiddle();
    w.debug_band = debug_band_;
    w.debug_cols = debug_cols_;
    FusedLaunch l = fp.launch;
    const int64_t units = (int64_t)n_channels * w.chunks_per_channel;
    l.grid = (int)std::min<int64_t>(units, resident);
    SYLDET_CUDA(launch_fused(l, fp.params, w, stream));
    launches_ += 1;
    return SYLDET_OK;
}

syldet_status Batch::launch_tc_range(const float *d_planar, int n_channels, int64_t n_samples, int64_t ch_stride, int64_t eval_offset,
                                     int64_t eval_count, int64_t evals_total, int detect_rule, float *d_all_outputs, EventSink sink,
                                     cudaStream_t stream, const int16_t *d_s16) {
    const Config &c = model_.config();
    const TcPlan &tp = model_.tc();
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return set_error(SYLDET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    // Y[channel][row][hop]: row r = samples [r*hop, (r+1)*hop) of a channel; only complete rows are part of the tensor
    const cuuint64_t dims[3] = {(cuuint64_t)c.hop, (cuuint64_t)(n_samples / c.hop), (cuuint64_t)n_channels};
    const cuuint64_t strides[2] = {(cuuint64_t)c.hop * 4, n_channels > 1 ? (cuuint64_t)ch_stride * 4 : (((cuuint64_t)n_samples * 4 + 15) & ~(cuuint64_t)15)};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t box_main[3] = {32, 64, 1}, box_tail[3] = {8, 64, 1};
    alignas(64) CUtensorMap tm_main, tm_tail;
    std::memset(&tm_main, 0, sizeof tm_main);
    std::memset(&tm_tail, 0, sizeof tm_tail);
    CUresult r1 = CUDA_SUCCESS, r2 = CUDA_SUCCESS;
    if (!d_s16) {   // the 16-bit source takes the direct data path: no tensor maps
    r1 = encode(&tm_main, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(d_planar), dims, strides, box_main, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    r2 = encode(&tm_tail, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(d_planar), dims, strides, box_tail, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS)
        return set_error(SYLDET_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r1) + ", " + std::to_string((int)r2) + ")");
    TcWork w{};
    w.n_channels = n_channels;
    w.evals_per_channel = eval_count;
    w.eval_offset = eval_offset;   // d_planar already points at the first sample of evaluation eval_offset
    w.out_evals_per_channel = evals_total;
    const int resident = model_.sm_count();
    // a unit covers whole tiles: chunk = n_tiles * tile_frames - (T - 1) evaluations (the first T-1 columns of a unit only warm
    // the window up); long enough that the warm-up is noise, short enough that every CTA gets many units
    const int64_t tf = tc_tile_frames(), warm = c.time_range - 1;
    static const int forced = [] { const char *e = std::getenv("SYLDET_TC_UNIT_TILES"); return e ? std::atoi(e) : 0; }();
    int64_t tiles_per_unit = tc_plan_unit_tiles(eval_count, n_channels, resident, (int)tf, (int)warm);
    if (forced > 0 && forced * tf > 4 * warm) tiles_per_unit = forced;
    const int64_t chunk = tiles_per_unit * tf - warm;
    w.chunk_evals = chunk;
    w.chunks_per_channel = (int)((eval_count + chunk - 1) / chunk);
    w.detect_rule = detect_rule;
    w.all_out = d_all_outputs;
    w.sink = sink;
    w.dft_hi = model_.dft_hi();
    w.dft_lo = model_.dft_lo();
    w.wcat_hi = model_.wcat_hi();
    w.wcat_lo = model_.wcat_lo();
    w.dft16 = model_.dft16();
    w.n0 = tp.n0;
    w.lo_stages = tc_lo_stages(tp.params, tp.hp);
    static const bool tf32_corr = std::getenv("SYLDET_TC_TF32_CORR") != nullptr;   // all three products in TF32 (amplitude-invariant)
    w.f16_corr = (tf32_corr || kernel_ == SYLDET_KERNEL_TENSOR_TF32 || !f16_ok_) ? 0 : 1;
    // int16-derived samples k / 32768 are exact fp16 operands (x: at most one rounding in the normal range; its tf32 residual times
    // 2^11 is a multiple of 2^-4 below 2): no lower bound needed there. Float input: see kTcGuardLo.
    w.guard_lo = pcm_exact_ ? 0.0f : kTcGuardLo;
    w.guard_range = pcm_exact_ ? 0.0f : kTcGuardRange;
    w.range_flag = reinterpret_cast<int *>(sink_count_.as<unsigned char>() + 8);
    w.debug_band = debug_band_;
    w.debug_cols = debug_cols_;
    // fp16 variant, experimental data path (SYLDET_TC_DIRECT=1): six splitter warps read the audio from global memory themselves -
    // no TMA, no raw fp32 tile in shared memory, 544 fewer shared-memory wavefronts per tile. Measured equal to slightly slower
    // than the TMA path (1.547 against 1.525 ms, profiles/r02_tc_direct_experiments.txt): a warp's loads share one scoreboard, so
    // the load latency of a tile is exposed once per tile. SYLDET_TC_PF: L2 prefetch distance in tiles for that path.
    static const int direct = [] { const char *e = std::getenv("SYLDET_TC_DIRECT"); return e ? std::atoi(e) : 0; }();
    static const int pf_dist = [] { const char *e = std::getenv("SYLDET_TC_PF"); return e ? std::atoi(e) : 0; }();
    w.direct = direct;
    w.pf_dist = std::max(0, std::min(pf_dist, 8));
    w.pcm = d_planar;
    w.pcm16 = d_s16;
    w.ch_stride = n_channels > 1 ? ch_stride : 0;
    w.n_rows = (int)std::min<int64_t>(n_samples / c.hop, INT32_MAX);
    const int64_t units = (int64_t)n_channels * w.chunks_per_channel;
    const int grid = (int)std::min<int64_t>(units, resident);
    static const bool timing = std::getenv("SYLDET_TC_TIMING") != nullptr;  // debug: per-role wait/busy cycles on stderr
    DeviceBuffer d_timing;
    if (timing) {
        syldet_status st = d_timing.reserve((size_t)grid * 32 * sizeof(long long));
        if (st != SYLDET_OK) return st;
        SYLDET_CUDA(cudaMemsetAsync(d_timing.get(), 0, (size_t)grid * 32 * sizeof(long long), stream));
        w.debug_timing = d_timing.as<long long>();
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (timing) {
        SYLDET_CUDA(cudaEventCreate(&ev0));
        SYLDET_CUDA(cudaEventCreate(&ev1));
        SYLDET_CUDA(cudaEventRecord(ev0, stream));
    }
    SYLDET_CUDA(launch_tc(tp.hp, grid, tp.smem, tp.params, w, &tm_main, &tm_tail, stream));
    launches_ += 1;
    if (timing) {
        std::vector<long long> h((size_t)grid * 32);
        SYLDET_CUDA(cudaEventRecord(ev1, stream));
        SYLDET_CUDA(cudaStreamSynchronize(stream));
        float ms = 0.0f;
        SYLDET_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        SYLDET_CUDA(cudaMemcpy(h.data(), d_timing.get(), h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        static const char *names[30] = {"tma.hi_free", "", "", "", "", "tma.total", "mma.full", "mma.tmem_empty", "mma.lo_ready", "mma.a_ready",
                                        "mma.p_empty", "mma.total", "F.p_full", "F.bar1", "F.bar2", "F.tmem_ld", "F.sums+l0", "F.total", "D.tmem_full", "",
                                        "D.a_free", "", "D.tmem_ld", "D.total", "S.full|ld", "S.lo_free", "S.convert", "", "", "S.total"};
        const double tiles = (double)n_channels * w.chunks_per_channel * ((w.chunk_evals + c.time_range - 1 + tc_tile_frames() - 1) / tc_tile_frames()) / grid;
        long long max_cycles = 0;   // slot 5 = whole role loop of the TMA warp
        for (int b = 0; b < grid; ++b) max_cycles = std::max(max_cycles, h[(size_t)b * 32 + 5]);
        std::fprintf(stderr, "[syldet tc timing] grid %d, ~%.0f tiles per CTA; launch %.3f ms, longest CTA %lld cycles => SM clock ~%.0f MHz; mean cycles per tile:\n",
                     grid, tiles, ms, max_cycles, ms > 0 ? (double)max_cycles / ms * 1e-3 : 0.0);
        for (int k = 0; k < 30; ++k) {
            if (!names[k][0]) continue;
            double sum = 0;
            for (int b = 0; b < grid; ++b) sum += (double)h[(size_t)b * 32 + k];
            std::fprintf(stderr, "  %-16s %10.0f\n", names[k], sum / grid / tiles);
        }
    }
    return SYLDET_OK;
}

// Wide-hidden path: per time segment, stft_planes_kernel (band magnitudes as A-operand planes + column statistics) and then
// wide_l0_kernel (3xTF32 tcgen05 contraction + the rest of the network + detection). The segment bounds the plane buffers.
syldet_status Batch::launch_wide_range(const float *d_planar, int n_channels, int64_t ch_stride, int64_t eval_begin, int64_t eval_count,
                                       int64_t evals_total, int detect_rule, float *d_all_outputs, EventSink sink, cudaStream_t stream) {
    const Config &c = model_.config();
    const WidePlan &wp = model_.wide();
    const int T = c.time_range, P = wp.params.n_planes, tile = wide_tile_rows();
    const int64_t budget = (int64_t)1 << 30;   // bytes of raw + lo planes per segment
    int64_t seg = std::max<int64_t>(tile, budget / ((int64_t)n_channels * P * 16 * 2));
    seg = (seg / tile) * tile;
    int seg_index = 0;
    for (int64_t e0 = eval_begin; e0 < eval_begin + eval_count; e0 += seg) {
        const int64_t ne = std::min(seg, eval_begin + eval_count - e0), ncols = ne + T - 1;
        const int64_t rows_alloc = ((ncols + kWideStftCols - 1) / kWideStftCols) * kWideStftCols + 4;
        syldet_status st = wide_hi_.reserve((size_t)n_channels * P * rows_alloc * 16);
        if (st != SYLDET_OK) return st;
        st = wide_lo_.reserve((size_t)n_channels * P * rows_alloc * 16);
        if (st != SYLDET_OK) return st;
        st = wide_stats_.reserve((size_t)n_channels * rows_alloc * sizeof(float4));
        if (st != SYLDET_OK) return st;
        // three timing events per segment (reused): stft_planes_kernel | wide_l0_kernel, read back by wide_phase_ms()
        while (wide_ev_.size() < (size_t)(3 * (seg_index + 1))) {
            cudaEvent_t e = nullptr;
            SYLDET_CUDA(cudaEventCreate(&e));
            wide_ev_.push_back(e);
        }
        SYLDET_CUDA(cudaEventRecord(wide_ev_[3 * seg_index], stream));
        static const bool ref_order_stft = std::getenv("SYLDET_WIDE_REF_STFT") != nullptr;   // reference-order radix-2 FFT instead of the Stockham kernel
        if (stft_planes_fast_supported(c.fourier_length) && !ref_order_stft)
            SYLDET_CUDA(launch_stft_planes_fast(model_.dev_net(), c.fourier_length, c.window_length, c.band, c.hop, d_planar, ch_stride, n_channels,
                                                e0, ncols, wide_hi_.as<float>(), wide_lo_.as<float>(), wide_stats_.as<float4>(), P, rows_alloc, stream));
        else
            SYLDET_CUDA(launch_stft_planes(model_.dev_net(), c.fourier_length, c.window_length, c.hop, d_planar, ch_stride, n_channels, e0, ncols,
                                           wide_hi_.as<float>(), wide_lo_.as<float>(), wide_stats_.as<float4>(), P, rows_alloc, stream));
        SYLDET_CUDA(cudaEventRecord(wide_ev_[3 * seg_index + 1], stream));
        if (debug_band_) {   // extractPower() values before the scaling: the same reference-order FFT, straight into the caller's buffer
            SYLDET_CUDA(launch_stft_band_generic(model_.dev_net(), c.fourier_length, d_planar, ch_stride, n_channels, e0, ncols,
                                                 debug_band_ + e0 * c.band, debug_cols_ * c.band, SYLDET_SCALING_LINEAR, stream));
            launches_ += 1;
        }
        WideWork w{};
        w.planes_hi = wide_hi_.as<float>();
        w.planes_lo = wide_lo_.as<float>();
        w.stats = wide_stats_.as<float4>();
        w.rows_alloc = rows_alloc;
        w.n_cols = ncols;
        w.n_evals = ne;
        w.n_channels = n_channels;
        w.detect_rule = detect_rule;
        w.eval_offset = e0;
        w.out_evals_per_channel = evals_total;
        w.all_out = d_all_outputs;
        w.sink = sink;
        w.weights = model_.wide_weights();
        w.v = model_.wide_v();
        w.bprime = model_.wide_bprime();
        w.w1 = model_.wide_w1();
        const int64_t tiles = (int64_t)n_channels * ((ne + tile - 1) / tile);
        SYLDET_CUDA(launch_wide((int)std::min<int64_t>(tiles, model_.sm_count()), wp.params, w, stream));
        SYLDET_CUDA(cudaEventRecord(wide_ev_[3 * seg_index + 2], stream));
        launches_ += 2;
        ++seg_index;
    }
    wide_segments_ = seg_index;
    return SYLDET_OK;
}

syldet_status Batch::wide_phase_ms(double *stft_ms, double *contraction_ms) {
    if (!last_.valid || wide_segments_ == 0) return set_error(SYLDET_ERR_ARG, "no wide-path launch to inspect");
    SYLDET_CUDA(cudaStreamSynchronize(last_.stream));
    double a = 0.0, b = 0.0;
    for (int k = 0; k < wide_segments_; ++k) {
        float ms = 0.f;
        SYLDET_CUDA(cudaEventElapsedTime(&ms, wide_ev_[3 * k], wide_ev_[3 * k + 1]));
        a += ms;
        SYLDET_CUDA(cudaEventElapsedTime(&ms, wide_ev_[3 * k + 1], wide_ev_[3 * k + 2]));
        b += ms;
    }
    if (stft_ms) *stft_ms = a;
    if (contraction_ms) *contraction_ms = b;
    return SYLDET_OK;
}

syldet_status Batch::launch_planar(const float *d_planar, int n_channels, int64_t n_samples, int64_t ch_stride,
                                   const float *valid_begin, const float *valid_end, int detect_rule, float *d_all_outputs,
                                   cudaStream_t stream) {
    const int64_t E = model_.config().num_evals(n_samples);
    return launch_planar_range(d_planar, n_channels, n_samples, n_samples, ch_stride, valid_begin, valid_end, 0, E, detect_rule,
                               d_all_outputs, true, stream);
}

// Evaluations [eval_begin, eval_begin + eval_count) of a planar buffer whose first n_avail samples per channel are present
// (n_samples is the length of the whole recording: it fixes the evaluation count and the pitch of the dense outputs).
syldet_status Batch::launch_planar_range(const float *d_planar, int n_channels, int64_t n_samples, int64_t n_avail, int64_t ch_stride,
                                         const float *valid_begin, const float *valid_end, int64_t eval_begin, int64_t eval_count,
                                         int detect_rule, float *d_all_outputs, bool reset_sink, cudaStream_t stream) {
    const Config &c = model_.config();
    const int64_t E = c.num_evals(n_samples);
    if (reset_sink) SYLDET_CUDA(cudaMemsetAsync(sink_count_.get(), 0, kSinkHeaderBytes, stream));
    if (eval_count <= 0) return SYLDET_OK;
    EventSink sink{sink_count_.as<unsigned long long>(), sink_events_.as<DevEvent>(), sink_outputs_.as<float>(), sink_capacity_};

    const int kernel = active_kernel();
    if (kernel == SYLDET_KERNEL_TENSOR || kernel == SYLDET_KERNEL_TENSOR_TF32) {
        // evaluations whose rows [e, e+T] are all complete hop-rows go to the tensor-core kernel; the last few (and inputs
        // whose base/pitch are not 16-byte aligned) go to the SIMT fused kernel
        const float *base = d_planar + eval_begin * c.hop;
        const int64_t n_rows = (n_avail - eval_begin * c.hop) / c.hop;
        const bool aligned = ((uintptr_t)base % 16 == 0) && (ch_stride % 4 == 0 || n_channels == 1);
        const int64_t e_tc = aligned ? std::max<int64_t>(0, std::min<int64_t>(eval_count, n_rows - c.time_range)) : 0;
        if (e_tc > 0) {
            syldet_status st = launch_tc_range(base, n_channels, n_avail - eval_begin * c.hop, ch_stride, eval_begin, e_tc, E, detect_rule,
                                               d_all_outputs, sink, stream);
            if (st != SYLDET_OK) return st;
        }
        if (e_tc < eval_count)
            return launch_fused_range(d_planar, n_channels, ch_stride, valid_begin, valid_end, eval_begin + e_tc, eval_count - e_tc, E,
                                      detect_rule, d_all_outputs, sink, stream);
        return SYLDET_OK;
    }
    if (kernel == SYLDET_KERNEL_FUSED)
        return launch_fused_range(d_planar, n_channels, ch_stride, valid_begin, valid_end, eval_begin, eval_count, E, detect_rule,
                                  d_all_outputs, sink, stream);
    if (kernel == SYLDET_KERNEL_WIDE)
        return launch_wide_range(d_planar, n_channels, ch_stride, eval_begin, eval_count, E, detect_rule, d_all_outputs, sink, stream);

    // generic two-kernel path, segmented in time so the band-feature buffer stays bounded
    const int L = c.band, T = c.time_range;
    const int64_t budget_cols = std::max<int64_t>(T + 1, (int64_t)(512ll << 20) / ((int64_t)n_channels * L * 4));
    const int64_t seg = std::max<int64_t>(1, budget_cols - (T - 1));
    for (int64_t e0 = eval_begin; e0 < eval_begin + eval_count; e0 += seg) {
        const int64_t ne = std::min(seg, eval_begin + eval_count - e0), ncols = ne + T - 1;
        syldet_status st = feat_.reserve((size_t)n_channels * ncols * L * sizeof(float));
        if (st != SYLDET_OK) return st;
        SYLDET_CUDA(launch_stft_band_generic(model_.dev_net(), c.fourier_length, d_planar, ch_stride, n_channels, e0, ncols,
                                             feat_.as<float>(), ncols * L, -1, stream));
        if (debug_band_) {   // extractPower() values before the scaling, straight into the caller's [channel][column][bin] buffer
            SYLDET_CUDA(launch_stft_band_generic(model_.dev_net(), c.fourier_length, d_planar, ch_stride, n_channels, e0, ncols,
                                                 debug_band_ + e0 * L, debug_cols_ * L, SYLDET_SCALING_LINEAR, stream));
            launches_ += 1;
        }
        SYLDET_CUDA(launch_nn_generic(model_.dev_net(), model_.max_width(), feat_.as<float>(), n_channels, ncols, ne, e0, E,
                                      detect_rule, d_all_outputs, sink, stream));
        launches_ += 2;
    }
    return SYLDET_OK;
}

syldet_status Batch::launch_device(const float *d_pcm, int n_channels, int64_t n_samples, int64_t ch_stride, int layout,
                                   int detect_rule, float *d_all_outputs, cudaStream_t stream, bool pcm_exact) {
    if (!d_pcm || n_channels <= 0 || n_samples < 0) return set_error(SYLDET_ERR_ARG, "bad pcm arguments");
    if (n_channels > 65535) return set_error(SYLDET_ERR_ARG, "more than 65535 channels in one call");
    if (layout == SYLDET_LAYOUT_PLANAR && n_channels > 1 && ch_stride < n_samples) return set_error(SYLDET_ERR_ARG, "channel_stride < n_samples");
    syldet_status st = use_device(model_.device());
    if (st != SYLDET_OK) return st;
    const Config &c = model_.config();
    const int64_t E = c.num_evals(n_samples);
    const unsigned long long total = (unsigned long long)std::max<int64_t>(E, 0) * n_channels;
    if (sink_capacity_ == 0 || !last_.valid || last_.n_channels != n_channels || last_.n_samples != n_samples) {
        st = ensure_sink(std::max<unsigned long long>(1, std::min<unsigned long long>(total, std::max<unsigned long long>(1ull << 20, total / 8))));
        if (st != SYLDET_OK) return st;
    }
    last_ = Last{true, d_pcm, n_channels, n_samples, ch_stride, layout, detect_rule, d_all_outputs, stream, pcm_exact};
    pcm_exact_ = pcm_exact;
    if (layout == SYLDET_LAYOUT_INTERLEAVED && n_channels > 1) {
        const int64_t pitch = (n_samples + 3) & ~(int64_t)3;
        st = planar_.reserve((size_t)n_channels * pitch * sizeof(float));
        if (st != SYLDET_OK) return st;
        SYLDET_CUDA(launch_ingest(d_pcm, SYLDET_PCM_F32, 1, n_channels, n_samples, 0, planar_.as<float>(), pitch, stream));
        launches_ += 1;
        return launch_planar(planar_.as<float>(), n_channels, n_samples, pitch, planar_.as<float>(),
                             planar_.as<float>() + (size_t)n_channels * pitch, detect_rule, d_all_outputs, stream);
    }
    const int64_t stride = n_channels > 1 ? ch_stride : n_samples;
    return launch_planar(d_pcm, n_channels, n_samples, stride, d_pcm, d_pcm + (size_t)(n_channels - 1) * stride + n_samples,
                         detect_rule, d_all_outputs, stream);
}

namespace {
// Detections leave the kernels in arrival order (one atomic per batch); the rows of the CLI are ordered by (channel, evaluation).
// key = channel << 40 | evaluation. LSD radix sort, 11 bits per pass over the digits that are in use: ~3 passes of sequential memory
// traffic instead of std::sort's ~20 compare-and-swap levels (400 000 detections per 1 h x 8 ch recording: 25 ms -> 3 ms).
void sort_event_keys(std::vector<EventKey> &keys) {
    const size_t n = keys.size();
    if (n < 2048) {
        std::sort(keys.begin(), keys.end(), [](const EventKey &a, const EventKey &b) { return a.key < b.key; });
        return;
    }
    uint64_t all = 0;
    for (const EventKey &k : keys) all |= k.key;
    static thread_local std::vector<EventKey> tmp;   // scratch kept between calls (no fresh page faults per recording)
    tmp.resize(n);
    std::vector<size_t> count(2049);
    EventKey *src = keys.data(), *dst = tmp.data();
    for (int shift = 0; shift < 64; shift += 11) {
        if (((all >> shift) & 0x7FF) == 0) continue;   // no key has a bit set in this digit: already ordered by it
        std::fill(count.begin(), count.end(), (size_t)0);
        for (size_t i = 0; i < n; ++i) ++count[((src[i].key >> shift) & 0x7FF) + 1];
        for (int d = 0; d < 2048; ++d) count[d + 1] += count[d];
        for (size_t i = 0; i < n; ++i) dst[count[(src[i].key >> shift) & 0x7FF]++] = src[i];
        std::swap(src, dst);
    }
    if (src != keys.data()) std::memcpy(keys.data(), src, n * sizeof(EventKey));
}
}  // namespace

// ---- ordering the detections on the device --------------------------------------------------------------------------------------
// Detections are unique per (channel, evaluation), so their order by (channel, evaluation) is a RANK, not a sort: mark bit
// channel * E + evaluation of a bitmap, prefix-sum the population counts, and detection i belongs at row
// popcount(bits before its own). Three small launches over n detections and C E / 32 words (1 h x 8 ch: 400 000 detections, 300 000
// words: tens of microseconds) replace the host's LSD radix sort (11 ms for the same recording: three scatter passes over 5 MB).
namespace {
constexpr int kOrderBlock = 1024;   // bitmap words per block of the prefix sum

__global__ void order_mark_kernel(const DevEvent *__restrict__ ev, unsigned long long n, long long evals, unsigned *__restrict__ bits) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long pos = (unsigned long long)ev[i].channel * (unsigned long long)evals + (unsigned long long)ev[i].eval;
    atomicOr(&bits[pos >> 5], 1u << (pos & 31));
}
// per block of kOrderBlock words: exclusive prefix of the population counts inside the block, and the block's total
__global__ void __launch_bounds__(kOrderBlock) order_scan_kernel(const unsigned *__restrict__ bits, unsigned long long n_words,
                                                                 unsigned *__restrict__ prefix, unsigned *__restrict__ block_tot) {
    __shared__ unsigned warp_tot[kOrderBlock / 32];
    const unsigned long long w = (unsigned long long)blockIdx.x * kOrderBlock + threadIdx.x;
    const unsigned c = w < n_words ? __popc(bits[w]) : 0u;
    unsigned incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((threadIdx.x & 31) >= d) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned v = warp_tot[threadIdx.x], inc2 = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc2, d);
            if (threadIdx.x >= d) inc2 += t;
        }
        warp_tot[threadIdx.x] = inc2 - v;   // exclusive
        if (threadIdx.x == 31) block_tot[blockIdx.x] = inc2;
    }
    __syncthreads();
    if (w < n_words) prefix[w] = warp_tot[threadIdx.x >> 5] + incl - c;
}
// exclusive prefix of the block totals, in place (one block; the number of blocks is small); the grand total goes behind them
__global__ void __launch_bounds__(1024) order_blocks_kernel(unsigned *__restrict__ block_tot, unsigned n_blocks) {
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < n_blocks; base += 1024) {
        const unsigned i = base + threadIdx.x;
        const unsigned c = i < n_blocks ? block_tot[i] : 0u;
        unsigned incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned v = warp_tot[threadIdx.x], inc2 = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc2, d);
                if (threadIdx.x >= d) inc2 += t;
            }
            warp_tot[threadIdx.x] = inc2 - v;
        }
        __syncthreads();
        const unsigned excl = carry + warp_tot[threadIdx.x >> 5] + incl - c;
        if (i < n_blocks) block_tot[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_tot[n_blocks] = carry;
}
__global__ void order_place_kernel(const DevEvent *__restrict__ ev, const float *__restrict__ outs, unsigned long long n, long long evals,
                                   int n_out, const unsigned *__restrict__ bits, const unsigned *__restrict__ prefix,
                                   const unsigned *__restrict__ block_tot, DevEvent *__restrict__ ev_sorted, float *__restrict__ outs_sorted) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevEvent e = ev[i];
    const unsigned long long pos = (unsigned long long)e.channel * (unsigned long long)evals + (unsigned long long)e.eval;
    const unsigned long long w = pos >> 5;
    const unsigned r = block_tot[w / kOrderBlock] + prefix[w] + __popc(bits[w] & ((1u << (pos & 31)) - 1u));
    ev_sorted[r] = e;
    for (int o = 0; o < n_out; ++o) outs_sorted[(size_t)r * n_out + o] = outs[(size_t)i * n_out + o];
}
}  // namespace

// Orders the n detections of the sink into sorted_events_ / sorted_outputs_ on `stream`. *ordered = false when the device path does
// not apply (bitmap above 1 GiB, or two detections claiming one (channel, evaluation): the caller then sorts on the host).
syldet_status Batch::order_events_on_device(unsigned long long n, int64_t evals, int n_channels, cudaStream_t stream, bool *ordered) {
    *ordered = false;
    const int O = model_.config().outputs;
    const unsigned long long positions = (unsigned long long)n_channels * (unsigned long long)std::max<int64_t>(evals, 1);
    const unsigned long long n_words = (positions + 31) / 32;
    if (n == 0 || n_words > (1ull << 28) || n > 0xFFFFFFF0ull) return SYLDET_OK;
    const unsigned n_blocks = (unsigned)((n_words + kOrderBlock - 1) / kOrderBlock);
    syldet_status st = order_bits_.reserve(n_words * 4);
    if (st == SYLDET_OK) st = order_prefix_.reserve(n_words * 4);
    if (st == SYLDET_OK) st = order_blocks_.reserve(((size_t)n_blocks + 1) * 4);
    // sized like the sink itself: a launch with a few more detections than any before must not re-allocate (cudaFree synchronises)
    const size_t cap = (size_t)std::max<unsigned long long>(n, sink_capacity_);
    if (st == SYLDET_OK) st = sorted_events_.reserve(cap * sizeof(DevEvent));
    if (st == SYLDET_OK) st = sorted_outputs_.reserve(cap * O * sizeof(float));
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaMemsetAsync(order_bits_.get(), 0, n_words * 4, stream));
    const unsigned ev_blocks = (unsigned)((n + 255) / 256);
    order_mark_kernel<<<ev_blocks, 256, 0, stream>>>(sink_events_.as<DevEvent>(), n, (long long)evals, order_bits_.as<unsigned>());
    order_scan_kernel<<<n_blocks, kOrderBlock, 0, stream>>>(order_bits_.as<unsigned>(), n_words, order_prefix_.as<unsigned>(), order_blocks_.as<unsigned>());
    order_blocks_kernel<<<1, 1024, 0, stream>>>(order_blocks_.as<unsigned>(), n_blocks);
    order_place_kernel<<<ev_blocks, 256, 0, stream>>>(sink_events_.as<DevEvent>(), sink_outputs_.as<float>(), n, (long long)evals, O, order_bits_.as<unsigned>(),
                                                      order_prefix_.as<unsigned>(), order_blocks_.as<unsigned>(), sorted_events_.as<DevEvent>(),
                                                      sorted_outputs_.as<float>());
    SYLDET_CUDA(cudaGetLastError());
    launches_ += 4;
    unsigned total = 0;
    SYLDET_CUDA(cudaMemcpyAsync(&total, order_blocks_.as<unsigned>() + n_blocks, 4, cudaMemcpyDeviceToHost, stream));
    SYLDET_CUDA(cudaStreamSynchronize(stream));
    *ordered = total == n;   // every detection marked its own bit
    return SYLDET_OK;
}

// Waits for the last launch and makes its results final. Two things can ask for a repeat of the launch: more detections than the
// event buffer holds (grow it to the worst case), and the range flag of the tensor kernel's fp16 correction pass (audio outside the
// window in which that pass is at float32 level: this handle switches to the all-TF32 variant for good). Dense outputs the caller
// asked for are rewritten by the repeat, on the same stream.
syldet_status Batch::settle(unsigned long long *n_events) {
    if (!last_.valid) return set_error(SYLDET_ERR_ARG, "no launch to inspect");
    syldet_status st = use_device(model_.device());
    if (st != SYLDET_OK) return st;
    for (int attempt = 0; attempt < 3; ++attempt) {
        SYLDET_CUDA(cudaStreamSynchronize(last_.stream));
        SinkHeader h{};
        SYLDET_CUDA(cudaMemcpy(&h, sink_count_.get(), sizeof h, cudaMemcpyDeviceToHost));
        bool replay = false;
        if (h.range_flag != 0 && f16_ok_) {
            f16_ok_ = false;
            ++range_fallbacks_;
            replay = true;
        }
        if (h.count > sink_capacity_) {
            const unsigned long long total = (unsigned long long)model_.config().num_evals(last_.n_samples) * last_.n_channels;
            st = ensure_sink(total);
            if (st != SYLDET_OK) return st;
            replay = true;
        }
        if (!replay) {
            *n_events = h.count;
            return SYLDET_OK;
        }
        const Last l = last_;
        st = launch_device(l.d_pcm, l.n_channels, l.n_samples, l.ch_stride, l.layout, l.detect_rule, l.d_all_outputs, l.stream, l.pcm_exact);
        if (st != SYLDET_OK) return st;
    }
    return set_error(SYLDET_ERR_OVERFLOW, "event buffer overflow after replay");
}

syldet_status Batch::last_detection_count(int64_t *count) {
    unsigned long long n = 0;
    syldet_status st = settle(&n);
    if (st != SYLDET_OK) return st;
    *count = (int64_t)n;
    return SYLDET_OK;
}

namespace {
// SYLDET_E2E_TIMING=1: host-side phase times of run_host / collect on stderr (debugging aid, adds synchronisation points)
bool e2e_timing() {
    static const bool on = std::getenv("SYLDET_E2E_TIMING") != nullptr;
    return on;
}
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

syldet_status Batch::collect(int64_t debounce_frames, Events &out) {
    if (!last_.valid) return set_error(SYLDET_ERR_ARG, "collect without a launch");
    if (debounce_frames < 0) return set_error(SYLDET_ERR_ARG, "negative debounce");
    const Config &c = model_.config();
    const int O = c.outputs;
    const double t_c0 = now_ms();
    unsigned long long n = 0;
    syldet_status st = settle(&n);
    if (st != SYLDET_OK) return st;
    const double t_c1 = now_ms();
    // pinned staging (shared with run_host's pipeline): the read-back runs at PCIe speed and needs no fresh pageable buffers
    st = ensure_pipeline(1, (size_t)sink_capacity_ * (sizeof(DevEvent) + sizeof(float) * O));
    if (st != SYLDET_OK) return st;
    const DevEvent *ev = static_cast<const DevEvent *>(h_events_);
    const float *outs = reinterpret_cast<const float *>(static_cast<const DevEvent *>(h_events_) + sink_capacity_);
    // rows ordered by (channel, evaluation): ranked on the device (order_events_on_device), or - large bitmaps - sorted here
    static const bool host_sort = std::getenv("SYLDET_HOST_SORT") != nullptr;
    bool ordered = false;
    if (n && !host_sort) {
        st = order_events_on_device(n, c.num_evals(last_.n_samples), last_.n_channels, last_.stream, &ordered);
        if (st != SYLDET_OK) return st;
    }
    if (n) {
        SYLDET_CUDA(cudaMemcpyAsync(h_events_, ordered ? sorted_events_.get() : sink_events_.get(), n * sizeof(DevEvent), cudaMemcpyDeviceToHost, last_.stream));
        SYLDET_CUDA(cudaMemcpyAsync(const_cast<float *>(outs), ordered ? sorted_outputs_.get() : sink_outputs_.get(), (size_t)n * O * sizeof(float),
                                    cudaMemcpyDeviceToHost, last_.stream));
        SYLDET_CUDA(cudaStreamSynchronize(last_.stream));
    }
    const double t_c2 = now_ms();
    out.outputs_per_event = O;
    out.rows.resize(n);
    out.outputs.resize((size_t)n * O);
    const int64_t first = c.first_output_sample();
    if (ordered) {
        for (size_t r = 0; r < n; ++r) out.rows[r] = syldet_event{ev[r].channel, 0, first + (int64_t)c.hop * ev[r].eval};  // TrackDetector.swift:39-42,67-68
        if (n) std::memcpy(out.outputs.data(), outs, (size_t)n * O * sizeof(float));
    } else {
        std::vector<EventKey> &keys = collect_keys_;
        keys.resize(n);
        for (size_t i = 0; i < n; ++i) keys[i] = EventKey{((uint64_t)(uint32_t)ev[i].channel << 40) | (uint64_t)ev[i].eval, (uint32_t)i};
        sort_event_keys(keys);
        for (size_t r = 0; r < n; ++r) {
            const DevEvent &e = ev[keys[r].idx];
            out.rows[r] = syldet_event{e.channel, 0, first + (int64_t)c.hop * e.eval};  // TrackDetector.swift:39-42,67-68
            for (int o = 0; o < O; ++o) out.outputs[r * O + o] = outs[(size_t)keys[r].idx * O + o];
        }
    }
    const double t_c3 = now_ms();
    debounce_sorted(c, out.rows, out.outputs, O, debounce_frames);
    if (e2e_timing())
        std::fprintf(stderr, "[syldet e2e] collect: wait %.2f ms, event d2h %.2f ms (%llu events), sort+rows %.2f ms, debounce %.2f ms\n", t_c1 - t_c0,
                     t_c2 - t_c1, n, t_c3 - t_c2, now_ms() - t_c3);
    return SYLDET_OK;
}

constexpr int kMaxSlices = 16;

syldet_status Batch::ensure_pipeline(int slices, size_t event_bytes) {
    if (!copy_stream_) SYLDET_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    if (!d2h_stream_) SYLDET_CUDA(cudaStreamCreateWithFlags(&d2h_stream_, cudaStreamNonBlocking));
    while ((int)ev_copied_.size() < slices) {
        cudaEvent_t a = nullptr, b = nullptr;
        SYLDET_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        ev_copied_.push_back(a);
        SYLDET_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ev_done_.push_back(b);
    }
    if (!h_counts_) SYLDET_CUDA(cudaMallocHost(&h_counts_, kMaxSlices * sizeof(SinkHeader)));
    if (event_bytes > h_events_bytes_) {
        if (h_events_) cudaFreeHost(h_events_);
        h_events_ = nullptr;
        h_events_bytes_ = 0;
        SYLDET_CUDA(cudaMallocHost(&h_events_, event_bytes));
        h_events_bytes_ = event_bytes;
    }
    return SYLDET_OK;
}

// The host entry point (what TrackDetector.process + the main.swift loop do for every track, TrackDetector.swift:45-105).
// The recording is cut into time slices: slice k+1 crosses PCIe while slice k is detected and its events are read back and
// sorted on the host, so everything except the copy itself hides behind the copy.
syldet_status Batch::run_host(const void *pcm, int fmt, int n_channels, int64_t n_samples, int64_t ch_stride, int layout,
                              int64_t debounce_frames, int detect_rule, float *all_outputs, Events &out, int trace_format, void *trace) {
    if (!pcm || n_channels <= 0 || n_samples < 0) return set_error(SYLDET_ERR_ARG, "bad pcm arguments");
    if (trace && trace_format != SYLDET_PCM_F32 && trace_format != SYLDET_PCM_S16) return set_error(SYLDET_ERR_ARG, "unknown trace format");
    if (n_channels > 65535) return set_error(SYLDET_ERR_ARG, "more than 65535 channels in one call");
    if (fmt != SYLDET_PCM_F32 && fmt != SYLDET_PCM_S16 && fmt != SYLDET_PCM_S24) return set_error(SYLDET_ERR_ARG, "unknown pcm format");
    if (layout != SYLDET_LAYOUT_PLANAR && layout != SYLDET_LAYOUT_INTERLEAVED) return set_error(SYLDET_ERR_ARG, "unknown layout");
    if (layout == SYLDET_LAYOUT_PLANAR && n_channels > 1 && ch_stride < n_samples) return set_error(SYLDET_ERR_ARG, "channel_stride < n_samples");
    if (debounce_frames < 0) return set_error(SYLDET_ERR_ARG, "negative debounce");
    syldet_status st = use_device(model_.device());
    if (st != SYLDET_OK) return st;
    const Config &c = model_.config();
    const int O = c.outputs;
    const int64_t E = c.num_evals(n_samples);
    const int64_t pitch = (n_samples + 3) & ~(int64_t)3;
    st = planar_.reserve(std::max<size_t>(16, (size_t)n_channels * pitch * sizeof(float)));
    if (st != SYLDET_OK) return st;
    const size_t esz = fmt == SYLDET_PCM_S16 ? 2 : fmt == SYLDET_PCM_S24 ? 3 : 4;
    const bool direct = fmt == SYLDET_PCM_F32 && layout == SYLDET_LAYOUT_PLANAR;   // lands in the planar buffer as is
    const bool inter = layout == SYLDET_LAYOUT_INTERLEAVED;
    const int64_t src_stride = inter ? 0 : (n_channels > 1 ? ch_stride : n_samples);
    if (!direct && n_samples > 0) {
        const size_t src_elems = inter ? (size_t)n_channels * n_samples : (size_t)(n_channels - 1) * src_stride + n_samples;
        st = staging_.reserve(src_elems * esz);
        if (st != SYLDET_OK) return st;
    }
    const unsigned long long total = (unsigned long long)E * n_channels;
    if (sink_capacity_ == 0 || !last_.valid || last_.n_channels != n_channels || last_.n_samples != n_samples) {
        st = ensure_sink(std::max<unsigned long long>(1, std::min<unsigned long long>(total, std::max<unsigned long long>(1ull << 20, total / 8))));
        if (st != SYLDET_OK) return st;
    }
    DeviceBuffer d_outs;
    if ((all_outputs || trace) && E > 0) {
        st = d_outs.reserve((size_t)n_channels * E * O * sizeof(float));
        if (st != SYLDET_OK) return st;
    }
    float *d_all = (all_outputs || trace) && E > 0 ? d_outs.as<float>() : nullptr;

    // ---- slices: evaluations [eb[k], eb[k+1]) become launchable once samples [0, sb[k+1]) are on the device ----------------
    int K = (int)std::min<int64_t>(kMaxSlices, std::max<int64_t>(1, (int64_t)total / slice_evals_));
    while (K > 1 && E / K < 4 * (int64_t)c.time_range + 16) --K;   // a slice is at least a few feature windows long
    int64_t eb[kMaxSlices + 1], sb[kMaxSlices + 1];
    sb[0] = 0;
    for (int k = 0; k <= K; ++k) eb[k] = k == K ? E : ((E * k / K) & ~(int64_t)3);   // multiples of 4: 16-byte aligned slice starts
    for (int k = 1; k <= K; ++k) {
        // everything evaluations < eb[k] read, and their T+1 hop-rows complete (the tensor-core kernel's tile rule)
        const int64_t need = std::max<int64_t>(c.samples_for_evals(eb[k]), (eb[k] + c.time_range + 1) * (int64_t)c.hop);
        sb[k] = k == K ? n_samples : std::min<int64_t>(n_samples, (need + 3) & ~(int64_t)3);
    }
    st = ensure_pipeline(K, (size_t)sink_capacity_ * (sizeof(DevEvent) + sizeof(float) * O));
    if (st != SYLDET_OK) return st;
    cudaStream_t sx = own_stream_;
    pcm_exact_ = fmt == SYLDET_PCM_S16;   // k / 32768: exact operands for the tensor kernel's fp16 correction pass
    last_ = Last{true, planar_.as<float>(), n_channels, n_samples, pitch, SYLDET_LAYOUT_PLANAR, detect_rule, d_all, sx, pcm_exact_};
    const double t_start = now_ms();
    float *planar = planar_.as<float>();
    // Errors inside the pipeline: nothing may stay in flight that reads the caller's buffer when control returns.
    auto drain = [&](syldet_status status) {
        cudaStreamSynchronize(copy_stream_);
        cudaStreamSynchronize(sx);
        cudaGetLastError();
        return status;
    };
#define SYLDET_CUDA_DRAIN(expr)                                                        \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) return drain(::syldet::cuda_fail(_e, #expr));           \
    } while (0)
    // 16-bit planar PCM on the reference's sample shape: the tensor kernel reads the staged int16 samples itself (direct data path of
    // kernels_tc.cu, k / 32768 in the splitter role) - no ingest kernel, no float32 copy of the recording in HBM. Only the last few
    // evaluations of the recording, which the SIMT fused kernel finishes, get their samples converted. SYLDET_S16_INGEST=1 keeps the
    // two-kernel path.
    static const bool s16_ingest = std::getenv("SYLDET_S16_INGEST") != nullptr;
    static const bool tf32_env = std::getenv("SYLDET_TC_TF32_CORR") != nullptr;
    const bool s16_direct = !s16_ingest && !tf32_env && fmt == SYLDET_PCM_S16 && !inter && kernel_ != SYLDET_KERNEL_TENSOR_TF32 && f16_ok_ &&
                            active_kernel() == SYLDET_KERNEL_TENSOR && tc_direct_s16_supported(model_.tc().hp, model_.tc().params) &&
                            src_stride % 4 == 0 && c.gap == 0 && n_samples > 0;
    for (int k = 0; k < K; ++k) {
        const int64_t s0 = sb[k], ns = sb[k + 1] - sb[k];
        if (ns > 0) {
            // cudaMemcpy2D rejects pitches above cudaDevAttrMaxPitch (2 GiB - 1), which a recording of more than ~3.4 hours per
            // channel would exceed: beyond it, one plain copy per channel row
            if (inter) {
                SYLDET_CUDA_DRAIN(cudaMemcpyAsync((char *)staging_.get() + (size_t)s0 * n_channels * esz, (const char *)pcm + (size_t)s0 * n_channels * esz,
                                                  (size_t)ns * n_channels * esz, cudaMemcpyHostToDevice, copy_stream_));
            } else {
                char *dst0 = direct ? (char *)(planar + s0) : (char *)staging_.get() + (size_t)s0 * esz;
                const size_t dpitch = direct ? (size_t)pitch * 4 : (size_t)src_stride * esz, spitch = (size_t)src_stride * esz;
                if (n_channels > 1 && std::max(dpitch, spitch) <= max_pitch_) {   // one 2-D copy per slice while the pitches allow it
                    SYLDET_CUDA_DRAIN(cudaMemcpy2DAsync(dst0, dpitch, (const char *)pcm + (size_t)s0 * esz, spitch, (size_t)ns * esz, n_channels,
                                                        cudaMemcpyHostToDevice, copy_stream_));
                } else {
                    for (int ch = 0; ch < n_channels; ++ch)
                        SYLDET_CUDA_DRAIN(cudaMemcpyAsync(dst0 + (size_t)ch * dpitch, (const char *)pcm + ((size_t)ch * src_stride + s0) * esz,
                                                          (size_t)ns * esz, cudaMemcpyHostToDevice, copy_stream_));
                }
            }
        }
        SYLDET_CUDA_DRAIN(cudaEventRecord(ev_copied_[k], copy_stream_));
        SYLDET_CUDA_DRAIN(cudaStreamWaitEvent(sx, ev_copied_[k], 0));
        if (s16_direct) {
            const int64_t e0 = eb[k], ne = eb[k + 1] - eb[k], n_avail = sb[k + 1];
            if (k == 0) SYLDET_CUDA_DRAIN(cudaMemsetAsync(sink_count_.get(), 0, kSinkHeaderBytes, sx));
            if (ne > 0) {
                EventSink sink{sink_count_.as<unsigned long long>(), sink_events_.as<DevEvent>(), sink_outputs_.as<float>(), sink_capacity_};
                const int16_t *base16 = staging_.as<int16_t>() + e0 * c.hop;
                const int64_t n_rows = (n_avail - e0 * c.hop) / c.hop;
                const int64_t e_tc = std::max<int64_t>(0, std::min<int64_t>(ne, n_rows - c.time_range));
                if (e_tc > 0) {
                    st = launch_tc_range(nullptr, n_channels, n_avail - e0 * c.hop, src_stride, e0, e_tc, E, detect_rule, d_all, sink, sx, base16);
                    if (st != SYLDET_OK) return drain(st);
                }
                if (e_tc < ne) {   // the end of the recording: these evaluations' samples as float32, then the SIMT fused kernel
                    const int64_t lo = ((e0 + e_tc) * c.hop) & ~(int64_t)7;
                    SYLDET_CUDA_DRAIN(launch_ingest((const char *)staging_.get() + (size_t)lo * esz, fmt, 0, n_channels, n_avail - lo, src_stride, planar + lo,
                                                    pitch, sx));
                    launches_ += 1;
                    st = launch_fused_range(planar, n_channels, pitch, planar, planar + (size_t)n_channels * pitch, e0 + e_tc, ne - e_tc, E, detect_rule,
                                            d_all, sink, sx);
                    if (st != SYLDET_OK) return drain(st);
                }
            }
        } else {
        if (!direct && ns > 0) {
            const char *src = (const char *)staging_.get() + (inter ? (size_t)s0 * n_channels * esz : (size_t)s0 * esz);
            SYLDET_CUDA_DRAIN(launch_ingest(src, fmt, inter ? 1 : 0, n_channels, ns, src_stride, planar + s0, pitch, sx));
            launches_ += 1;
        }
        st = launch_planar_range(planar, n_channels, n_samples, sb[k + 1], pitch, planar, planar + (size_t)n_channels * pitch, eb[k],
                                 eb[k + 1] - eb[k], detect_rule, d_all, k == 0, sx);
        if (st != SYLDET_OK) return drain(st);
        }
        SYLDET_CUDA_DRAIN(cudaMemcpyAsync(h_counts_ + k, sink_count_.get(), sizeof(SinkHeader), cudaMemcpyDeviceToHost, sx));
        SYLDET_CUDA_DRAIN(cudaEventRecord(ev_done_[k], sx));
    }

    // ---- per slice: wait, read its events back, sort them by (channel, evaluation) while later slices are still in flight ----
    using Key = EventKey;
    DevEvent *h_ev = static_cast<DevEvent *>(h_events_);
    float *h_out = reinterpret_cast<float *>(h_ev + sink_capacity_);
    // the result vectors are touched now, while the copies run, so that the tail below does not pay their page faults
    out.rows.resize(last_event_count_);
    out.outputs.resize((size_t)last_event_count_ * O);
    // per slice: its rows in (channel, evaluation) order, built while later slices are still crossing PCIe; seg[k][ch] = first row of channel ch
    std::vector<std::vector<syldet_event>> slice_rows(K);
    std::vector<std::vector<float>> slice_outs(K);
    std::vector<std::vector<size_t>> seg(K, std::vector<size_t>((size_t)n_channels + 1, 0));
    std::vector<Key> keys;
    const int64_t first = c.first_output_sample();
    unsigned long long done = 0;
    bool redo = false;   // the event buffer overflowed or the fp16 range flag went up: collect() repeats the launch (see settle)
    double t_copied = 0.0;
    for (int k = 0; k < K; ++k) {
        SYLDET_CUDA_DRAIN(cudaEventSynchronize(ev_done_[k]));
        if (k == K - 1) t_copied = now_ms();
        const unsigned long long n_k = h_counts_[k].count;
        if (n_k > sink_capacity_ || (h_counts_[k].range_flag != 0 && f16_ok_)) {
            redo = true;
            break;
        }
        const unsigned long long m = n_k - done;
        if (m == 0) continue;
        SYLDET_CUDA_DRAIN(cudaMemcpyAsync(h_ev + done, sink_events_.as<DevEvent>() + done, m * sizeof(DevEvent), cudaMemcpyDeviceToHost, d2h_stream_));
        SYLDET_CUDA_DRAIN(cudaMemcpyAsync(h_out + done * O, sink_outputs_.as<float>() + done * O, m * O * sizeof(float), cudaMemcpyDeviceToHost,
                                          d2h_stream_));
        SYLDET_CUDA_DRAIN(cudaStreamSynchronize(d2h_stream_));
        keys.resize(m);
        for (unsigned long long i = 0; i < m; ++i) {
            const DevEvent &e = h_ev[done + i];
            keys[i] = Key{((uint64_t)(uint32_t)e.channel << 40) | (uint64_t)e.eval, (uint32_t)(done + i)};
        }
        sort_event_keys(keys);
        std::vector<syldet_event> &rows = slice_rows[k];
        std::vector<float> &outs = slice_outs[k];
        rows.resize(m);
        outs.resize((size_t)m * O);
        int ch_next = 0;
        for (size_t i = 0; i < m; ++i) {
            const DevEvent &e = h_ev[keys[i].idx];
            while (ch_next <= e.channel) seg[k][ch_next++] = i;
            rows[i] = syldet_event{e.channel, 0, first + (int64_t)c.hop * e.eval};  // TrackDetector.swift:39-42,67-68
            for (int o = 0; o < O; ++o) outs[i * O + o] = h_out[(size_t)keys[i].idx * O + o];
        }
        while (ch_next <= n_channels) seg[k][ch_next++] = m;
        done = n_k;
    }
    if (redo) {
        // the recording is resident once the pipeline has drained: collect() sees the same condition in the device-side header
        // and repeats the launch in one piece (larger event buffer / all-TF32 variant)
        SYLDET_CUDA(cudaStreamSynchronize(copy_stream_));
        SYLDET_CUDA(cudaStreamSynchronize(sx));
        if (s16_direct) {   // the repeat runs from the float32 copy of the recording, which this path had not made
            SYLDET_CUDA(launch_ingest(staging_.get(), fmt, 0, n_channels, n_samples, src_stride, planar, pitch, sx));
            launches_ += 1;
        }
        st = collect(debounce_frames, out);
        if (st != SYLDET_OK) return st;
    } else {
        // ---- merge: slices are consecutive in time, so per channel the slices' segments simply follow each other (block copies) ----
        const double t_m0 = now_ms();
        out.outputs_per_event = O;
        out.rows.resize(done);
        out.outputs.resize((size_t)done * O);
        last_event_count_ = (size_t)done + (size_t)done / 16;
        size_t r = 0;
        for (int ch = 0; ch < n_channels; ++ch)
            for (int k = 0; k < K; ++k) {
                const size_t a0 = seg[k][ch], a1 = seg[k][ch + 1];
                if (a1 > a0) {
                    std::memcpy(out.rows.data() + r, slice_rows[k].data() + a0, (a1 - a0) * sizeof(syldet_event));
                    std::memcpy(out.outputs.data() + r * O, slice_outs[k].data() + a0 * O, (a1 - a0) * O * sizeof(float));
                    r += a1 - a0;
                }
            }
        const double t_m1 = now_ms();
        debounce_sorted(c, out.rows, out.outputs, O, debounce_frames);
        if (e2e_timing())
            std::fprintf(stderr, "[syldet e2e] %d slices: last slice done %.2f ms after start; tail: sort %.2f ms, merge %.2f ms, debounce %.2f ms (%llu events)\n",
                         K, t_copied - t_start, t_m0 - t_copied, t_m1 - t_m0, now_ms() - t_m1, done);
    }
    if (all_outputs && E > 0) {
        SYLDET_CUDA(cudaStreamSynchronize(sx));
        SYLDET_CUDA(cudaMemcpy(all_outputs, d_outs.get(), (size_t)n_channels * E * O * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (trace && n_samples > 0) {  // simulator trace (ViewControllerSimulator.swift:251-254, 308-344), built on the device
        const size_t tsz = trace_format == SYLDET_PCM_S16 ? 2 : 4;
        DeviceBuffer d_trace;
        st = d_trace.reserve((size_t)n_channels * n_samples * tsz);
        if (st != SYLDET_OK) return st;
        SYLDET_CUDA(launch_simulator_trace(d_all, n_channels, E, O, (float)c.thresholds[0], c.first_output_sample(), c.hop, n_samples,
                                           trace_format, d_trace.get(), n_samples, sx));
        launches_ += 1;
        SYLDET_CUDA(cudaMemcpyAsync(trace, d_trace.get(), (size_t)n_channels * n_samples * tsz, cudaMemcpyDeviceToHost, sx));
        SYLDET_CUDA(cudaStreamSynchronize(sx));
    }
    return SYLDET_OK;
#undef SYLDET_CUDA_DRAIN
}

// extractPower()[f0 ..< f1] (CSTFT.swift:280-337, SyllableDetector.swift:134-151) as the active kernel computes it: the detection
// path runs with a tap on its band magnitudes (before the spectrogram scaling). Test / inspection entry point, not a fast path.
syldet_status Batch::spectra_host(const void *pcm, int fmt, int n_channels, int64_t n_samples, int64_t ch_stride, int layout, float *band,
                                  int64_t *n_columns) {
    const Config &c = model_.config();
    const int64_t E = c.num_evals(n_samples);
    const int64_t cols = E > 0 ? E + c.time_range - 1 : 0;
    if (n_columns) *n_columns = cols;
    if (!band || cols == 0) return SYLDET_OK;   // query of *n_columns, or nothing to compute
    syldet_status st = use_device(model_.device());
    if (st != SYLDET_OK) return st;
    DeviceBuffer d_band;
    const size_t bytes = (size_t)n_channels * cols * c.band * sizeof(float);
    st = d_band.reserve(bytes);
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaMemset(d_band.get(), 0, bytes));
    debug_band_ = d_band.as<float>();
    debug_cols_ = cols;
    Events ev;
    st = run_host(pcm, fmt, n_channels, n_samples, ch_stride, layout, 0, SYLDET_DETECT_ANY_OUTPUT, nullptr, ev);
    debug_band_ = nullptr;
    debug_cols_ = 0;
    if (st != SYLDET_OK) return st;
    SYLDET_CUDA(cudaStreamSynchronize(own_stream_));
    SYLDET_CUDA(cudaMemcpy(band, d_band.get(), bytes, cudaMemcpyDeviceToHost));
    return SYLDET_OK;
}

}  // namespace syldet
