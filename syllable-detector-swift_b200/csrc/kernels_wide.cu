// Wide-hidden tensor path (sm_100a): NeuralNetLayer.apply of layer 0 (Common/NeuralNet.swift:366-377) for networks whose hidden
// layer makes it a real dense contraction - BASELINE config 4: FFT 1024, hop 4, 162 bins x 8 columns = 1296 inputs -> 256..1024
// hidden units - as a 3xTF32 tcgen05 GEMM with the accumulator in TMEM, followed in the same kernel by the rest of
// NeuralNet.apply (:294-326: window normaliser folded into a per-row scale / offset, transfer function, output layer, reverse output
// maps) and the threshold test of TrackDetector.process (SyllableDetectorCLI/TrackDetector.swift:71-77).
//
//   Z[j][h] = sum_{t < T} sum_{f < L} m[j + t][f] * W'[h][t*L + f]          m = band magnitudes (stft_planes_kernel), W' = folded weights
//
// The A operand is never materialised: row j of the [evaluations x T*L] input matrix is T consecutive rows of the magnitude matrix
// (the sliding feature window of SyllableDetector.processNewValue, Common/SyllableDetector.swift:153-217). The magnitudes sit in
// shared memory as planes of 4 bins, [plane][row][16 B] - the canonical no-swizzle K-major UMMA layout (core matrix = 8 rows x 16 B,
// contiguous) - so the operand of column offset t is the SAME bytes with the start address advanced by t rows (16 t bytes): an
// implicit im2col along time at zero cost. K runs over (chunk of kWidePL planes, t, plane); the magnitudes of a chunk are loaded once
// per tile and used by all T offsets, the weights stream through a four-stage ring of pre-arranged blocks (cp.async.bulk, one per plane).
//   per tile: 256 evaluations (two M = 128 accumulators of N = 256 columns: all 512 TMEM columns), so that every weight block that
//   crosses L2 -> shared memory feeds 2 x 128 rows (the weight stream is what bounds this kernel after the tensor pipe).
// Roles: warp 0 loader (cp.async.bulk), warp 1 MMA issuer + TMEM, warps 2-9 epilogue (one per TMEM lane quadrant and accumulator; a
// thread owns one accumulator row = one evaluation, so the output layer needs no cross-lane reduction).
#include <cuda.h>

#include "fft_regs.cuh"
#include "fused_epilogue.cuh"
#include "ptx_sm100.cuh"

namespace syldet {

namespace {

constexpr int kWThreads = 10 * 32;                  // loader, MMA issuer, 2 x 4 epilogue warps
constexpr int kWTileRows = 256;                     // evaluations per tile (2 accumulators x 128 TMEM lanes)
constexpr int kWRowsPad = kWTileRows + 16;          // magnitude rows in shared memory: tile + T - 1 (T <= 17)
constexpr int kWN = 256;                            // hidden units per accumulator pass
constexpr int kWPlaneA = kWRowsPad * 16;            // bytes of one A plane
constexpr int kWPlaneW = kWN * 16;                  // bytes of one weight plane
constexpr int kWABytes = 2 * kWidePL * kWPlaneA;    // one A buffer: raw | lo
constexpr int kWWBytes = 2 * kWidePL * kWPlaneW;    // one weight block of a (chunk, t) step: hi | lo
constexpr int kWStageBytes = kWidePL * kWPlaneW;    // one stage of the weight ring: the hi OR the lo part of a block
constexpr int kWStages = 4;                         // ring depth: three stages (72 KB) can be in flight while one is consumed
constexpr int kWTmemCols = 512;

struct WideSmem {
    static constexpr int a0 = 0, a1 = kWABytes, w0 = 2 * kWABytes, cst = w0 + kWStages * kWStageBytes;   // then float4 cst[h_pad], barriers
    __host__ __device__ static constexpr int bars(int h_pad) { return cst + h_pad * 16; }
    __host__ __device__ static constexpr int total(int h_pad) { return bars(h_pad) + 256; }
};
static_assert(kWABytes % 128 == 0 && kWStageBytes % 128 == 0, "alignment of the operand buffers");

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units):
// core matrix = 8 rows x 16 B contiguous; lbo = bytes between the two core matrices of a K step, sbo = bytes between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}

// The rest of NeuralNet.apply for the rows of one accumulator (NeuralNet.swift:310-323): z = acc * alpha + beta * V + B', hidden
// transfer function, output layer, output transfer, reverse maps in index order, threshold test (TrackDetector.swift:71-77), events.
// cst[h] = {V_h, B'_h, W1[0][h], W1[1][h]}: outputs 0 and 1 come from shared memory, further outputs (rare) from global memory.
template <int TF0>
__device__ __forceinline__ void wide_epilogue(const WideParams &p, const WideWork &w, const float4 *__restrict__ cst, uint64_t *acc_full,
                                              uint64_t *acc_empty, uint32_t tmem_base, int warp, int lane, int n_tiles, int tiles_per_ch, int n_nc) {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + half * kWN;
    const int n_out = p.n_out, T = p.time_range;
    uint32_t acc_use = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int ch = tile / tiles_per_ch;
        const int64_t j = (int64_t)(tile - ch * tiles_per_ch) * kWTileRows + half * 128 + quad * 32 + lane;
        const bool valid = j < w.n_evals;
        float inv = 1.0f, beta = 0.0f;
        if (valid && p.window_stat != FUSED_STAT_NONE) {      // window statistic from the per-column partials
            const float4 *st = w.stats + (int64_t)ch * w.rows_alloc + j;
            float s0 = p.window_stat == FUSED_STAT_L2 ? 0.0f : INFINITY, s1 = -INFINITY;
            for (int t = 0; t < T; ++t) {
                const float4 c4 = __ldg(st + t);
                if (p.window_stat == FUSED_STAT_L2) s0 += c4.x;
                else { s0 = fminf(s0, c4.y); s1 = fmaxf(s1, c4.z); }
            }
            if (p.window_stat == FUSED_STAT_L2) inv = rcp_fast(sqrt_fast(s0));   // silence: 0 * inf = NaN (NeuralNet.swift:47-59)
            else {
                const float range = s1 - s0;
                if (0 == range) { inv = 0.0f; beta = -1.0f; }                     // flat window (NeuralNet.swift:84-88)
                else { inv = 2.0f / range; beta = (0 - s0 - s1) / range; }
            }
        }
        float out[kFusedMaxOut];
#pragma unroll
        for (int o = 0; o < kFusedMaxOut; ++o) out[o] = o < n_out ? p.b1[o] : 0.0f;
        for (int nc = 0; nc < n_nc; ++nc, ++acc_use) {
            ptx::mbar_wait(acc_full, acc_use & 1);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < kWN / 32; ++cb) {
                uint32_t r[32];
                ptx::tmem_ld_x32(taddr + cb * 32, r);
                ptx::tc_wait_ld();
                const int h0 = nc * kWN + cb * 32;
                if (h0 < p.hidden) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float4 c4 = cst[h0 + i];
                        const float a = transfer_fast(TF0, fmaf(__uint_as_float(r[i]), inv, fmaf(beta, c4.x, c4.y)));
                        out[0] = fmaf(c4.z, a, out[0]);
                        out[1] = fmaf(c4.w, a, out[1]);
                        if (n_out > 2) {
                            const float *w1 = w.w1 + h0 + i;                 // [n_out][h_pad]; padded hidden units have zero weights
#pragma unroll
                            for (int o = 2; o < kFusedMaxOut; ++o)
                                if (o < n_out) out[o] = fmaf(__ldg(w1 + (size_t)o * p.h_pad), a, out[o]);
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acc_empty);
        }
        // output transfer, reverse maps in index order (NeuralNet.swift:316-323), threshold test, events
        bool hit = false;
#pragma unroll
        for (int o = 0; o < kFusedMaxOut; ++o) {
            if (o < n_out) {
                float v = transfer_fast(p.tf1, out[o]);
                for (int k = 0; k < p.n_op; ++k) v = (v + (0 - p.op_y[k])) / p.op_gain[k * kFusedMaxOut + o] + p.op_xoff[k * kFusedMaxOut + o];
                out[o] = v;
                if (v >= p.thr_f[o] && (w.detect_rule == SYLDET_DETECT_ANY_OUTPUT || o == 0)) hit = true;   // NaN -> false
            }
        }
        hit = hit && valid;
        if (valid && w.all_out) {
            float *dst = w.all_out + ((int64_t)ch * w.out_evals_per_channel + w.eval_offset + j) * n_out;
#pragma unroll
            for (int o = 0; o < kFusedMaxOut; ++o)
                if (o < n_out) dst[o] = out[o];
        }
        const unsigned hits = __ballot_sync(0xffffffffu, hit);
        if (hits) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(w.sink.count, (unsigned long long)__popc(hits));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
                const unsigned long long idx = base + __popc(hits & ((1u << lane) - 1));
                if (idx < w.sink.capacity) {
                    w.sink.events[idx] = DevEvent{ch, 0, w.eval_offset + j};
#pragma unroll
                    for (int o = 0; o < kFusedMaxOut; ++o)
                        if (o < n_out) w.sink.outputs[idx * n_out + o] = out[o];
                }
            }
        }
    }
}

__device__ __forceinline__ float tf32_trunc_w(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(kWThreads, 1) wide_l0_kernel(const __grid_constant__ WideParams p, const WideWork w) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char *smem = smem_dyn + ((128u - (ptx::smem_addr(smem_dyn) & 127u)) & 127u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.time_range, n_nc = p.h_pad / kWN, n_chunks = p.n_planes / kWidePL;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + WideSmem::bars(p.h_pad));
    uint64_t *a_full = bars, *a_empty = bars + 2, *w_full = bars + 4, *w_empty = bars + 4 + kWStages, *acc_full = bars + 4 + 2 * kWStages,
             *acc_empty = acc_full + 1;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 1);
    float4 *cst = reinterpret_cast<float4 *>(smem + WideSmem::cst);   // {V_h, B'_h, W1[0][h], W1[1][h]}

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&a_full[i], 1);
            ptx::mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < kWStages; ++i) {
            ptx::mbar_init(&w_full[i], 1);
            ptx::mbar_init(&w_empty[i], 1);
        }
        ptx::mbar_init(acc_full, 1);
        ptx::mbar_init(acc_empty, 8);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_ptr, kWTmemCols);
        ptx::tmem_relinquish();
    }
    for (int h = tid; h < p.h_pad; h += kWThreads)
        cst[h] = make_float4(__ldg(w.v + h), __ldg(w.bprime + h), __ldg(w.w1 + h), p.n_out > 1 ? __ldg(w.w1 + p.h_pad + h) : 0.0f);
    // rows of the A buffers that no copy fills (the tail of the last tile of a channel) must hold finite values
    for (int i = tid; i < 2 * kWABytes / 16; i += kWThreads) reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int tiles_per_ch = (int)((w.n_evals + kWTileRows - 1) / kWTileRows);
    const int n_tiles = tiles_per_ch * w.n_channels;

    if (warp == 0) {
        // ================================ loader ======================================================================
        if (ptx::elect_one()) {
            uint32_t a_use = 0, w_use = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int ch = tile / tiles_per_ch;
                const int64_t j0 = (int64_t)(tile - ch * tiles_per_ch) * kWTileRows;
                const int64_t rows_left = w.n_cols - j0;                       // magnitude rows that exist from j0 on
                const int rows = (int)min((int64_t)(kWTileRows + T - 1), rows_left);
                const float *hi = w.planes_hi + ((int64_t)ch * p.n_planes * w.rows_alloc + j0) * 4;
                const float *lo = w.planes_lo + ((int64_t)ch * p.n_planes * w.rows_alloc + j0) * 4;
                for (int nc = 0; nc < n_nc; ++nc) {
                    for (int c = 0; c < n_chunks; ++c, ++a_use) {
                        const int ab = a_use & 1;
                        ptx::mbar_wait(&a_empty[ab], ((a_use >> 1) & 1) ^ 1);
                        unsigned char *dst = smem + (ab ? WideSmem::a1 : WideSmem::a0);
                        ptx::mbar_expect_tx(&a_full[ab], 2u * kWidePL * (uint32_t)rows * 16u);
                        for (int q = 0; q < kWidePL; ++q) {
                            const int64_t off = (int64_t)(c * kWidePL + q) * w.rows_alloc * 4;
                            ptx::bulk_copy_g2s(dst + q * kWPlaneA, hi + off, (uint32_t)rows * 16u, &a_full[ab]);
                            ptx::bulk_copy_g2s(dst + (kWidePL + q) * kWPlaneA, lo + off, (uint32_t)rows * 16u, &a_full[ab]);
                        }
                        for (int t = 0; t < T; ++t) {
                            const unsigned char *blk = reinterpret_cast<const unsigned char *>(w.weights) +
                                                       (size_t)((nc * n_chunks + c) * T + t) * kWWBytes;
                            for (int part = 0; part < 2; ++part, ++w_use) {      // hi, then lo: one ring stage each
                                const int ws = w_use % kWStages;
                                ptx::mbar_wait(&w_empty[ws], ((w_use / kWStages) & 1) ^ 1);
                                ptx::mbar_expect_tx(&w_full[ws], kWStageBytes);
                                unsigned char *dst = smem + WideSmem::w0 + ws * kWStageBytes;
                                // one copy per plane: several requests in flight instead of one long one
                                for (int q = 0; q < kWidePL; ++q)
                                    ptx::bulk_copy_g2s(dst + q * kWPlaneW, blk + part * kWStageBytes + q * kWPlaneW, kWPlaneW, &w_full[ws]);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================================================
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::idesc_tf32(128, kWN);
            const uint32_t a_base0 = ptx::smem_addr(smem + WideSmem::a0), a_base1 = ptx::smem_addr(smem + WideSmem::a1);
            const uint32_t w_base0 = ptx::smem_addr(smem + WideSmem::w0);
            uint32_t a_use = 0, w_use = 0, acc_use = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int nc = 0; nc < n_nc; ++nc, ++acc_use) {
                    ptx::mbar_wait(acc_empty, (acc_use & 1) ^ 1);       // the epilogue has read the previous accumulators
                    ptx::tc_fence_after();
                    uint32_t acc = 0;
                    for (int c = 0; c < n_chunks; ++c, ++a_use) {
                        const int ab = a_use & 1;
                        ptx::mbar_wait(&a_full[ab], (a_use >> 1) & 1);
                        for (int t = 0; t < T; ++t) {
#pragma unroll 1
                            for (int part = 0; part < 2; ++part, ++w_use) {   // weight hi part: raw x hi, lo x hi; weight lo part: raw x lo   (3xTF32)
                                const int ws = w_use % kWStages;
                                ptx::mbar_wait(&w_full[ws], (w_use / kWStages) & 1);
                                ptx::tc_fence_after();
                                const uint32_t w_part = w_base0 + ws * kWStageBytes;
#pragma unroll 1
                                for (int half = 0; half < 2; ++half) {
                                    const uint32_t d = tmem_base + half * kWN;
                                    const uint32_t a_row = (ab ? a_base1 : a_base0) + (uint32_t)(half * 128 + t) * 16u;
                                    uint32_t accf = part ? 1u : acc;
#pragma unroll 1
                                    for (int pass = 0; pass < (part ? 1 : 2); ++pass) {
                                        const uint32_t a_part = a_row + (pass == 1 ? kWidePL * kWPlaneA : 0);
#pragma unroll
                                        for (int ks = 0; ks < kWidePL / 2; ++ks) {
                                            ptx::mma_tf32_ss(d, smem_desc_nosw(a_part + 2 * ks * kWPlaneA, kWPlaneA, 128),
                                                             smem_desc_nosw(w_part + 2 * ks * kWPlaneW, kWPlaneW, 128), idesc, accf);
                                            accf = 1;
                                        }
                                    }
                                }
                                ptx::mma_commit(&w_empty[ws]);
                            }
                            acc = 1;
                        }
                        ptx::mma_commit(&a_empty[ab]);
                    }
                    ptx::mma_commit(acc_full);
                }
            }
        }
    } else {
        // ================================ epilogue ====================================================================
        // Eight warps: warps 2-5 read accumulator 0, warps 6-9 accumulator 1 (TMEM lane quadrant = warp % 4). A thread owns one row =
        // one evaluation. The hidden transfer function is a compile-time parameter of the loop body (one dispatch per kernel).
        switch (p.tf0) {
            case SYLDET_TF_TANSIG: wide_epilogue<SYLDET_TF_TANSIG>(p, w, cst, acc_full, acc_empty, tmem_base, warp, lane, n_tiles, tiles_per_ch, n_nc); break;
            case SYLDET_TF_LOGSIG: wide_epilogue<SYLDET_TF_LOGSIG>(p, w, cst, acc_full, acc_empty, tmem_base, warp, lane, n_tiles, tiles_per_ch, n_nc); break;
            case SYLDET_TF_SATLIN: wide_epilogue<SYLDET_TF_SATLIN>(p, w, cst, acc_full, acc_empty, tmem_base, warp, lane, n_tiles, tiles_per_ch, n_nc); break;
            default: wide_epilogue<SYLDET_TF_PURELIN>(p, w, cst, acc_full, acc_empty, tmem_base, warp, lane, n_tiles, tiles_per_ch, n_nc); break;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, kWTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// stft_planes_fast_kernel: CircularShortTimeFourierTransform.extractPower (CSTFT.swift:280-337) for the high-overlap shape, band
// slice + spectrogram scaling of SyllableDetector.processFourierData / processNewValue (Common/SyllableDetector.swift:134-212).
// A CTA stages the audio span of kWideStftCols consecutive columns once (hop 4, window 1024: 64 columns share 1276 samples instead
// of reading 65 536); each warp then transforms one frame at a time: the real N-point FFT as an M = N/2-point complex Stockham
// autosort transform in shared memory, radix-8 stages (a last radix-2 / radix-4 stage when log2 M is not a multiple of 3), the
// butterflies in registers (fft_regs.cuh), ping-pong buffers XOR-swizzled (padf) so that every stage's stride-8^s stores and
// unit-stride loads are bank-conflict free. The band bins are untangled from the packed transform, scaled, and the tile
// leaves the CTA in the plane layout wide_l0_kernel consumes ([plane][row][4 bins], raw and v - tf32(v)) together with the
// per-column statistics of the per-window normalisers. (stft_planes_kernel in kernels_generic.cu is the reference-order fallback.)
// Position of element i of a ping-pong buffer: the low four bits (the bank pair of a float2) are XOR-swizzled with bits 4, 5 and 6,
//   bank bits = (b0 ^ b4, b1 ^ b5, b2 ^ b6, b3 ^ b6),
// which makes every access pattern of a half warp (16 lanes x 8 bytes = one wavefront) a bijection onto the 16 bank pairs: the
// unit-stride loads (index bits 0-3 vary), the stores of the first stage (stride 8: bits 3-6), of the second (bits 0-2 and 6) and of
// the later ones (bits 0-3). Padding by one element per eight (round 2, first version) left one two-way conflict in every
// unit-stride load and in the second stage's stores: 256 wavefronts per 512-point frame where 160 are needed
// (profiles/r02_wide_v2_ncu_summary.txt: 116 M conflicts).
__device__ __forceinline__ int padf(int i) { return i ^ ((i >> 4) & 3) ^ (((i >> 6) & 1) * 12); }

// One Stockham stage of radix R over the M-point sequence: out[(j / Ns) Ns R + j % Ns + r Ns] = DFT_R(in[j + r M / R] w^{r (j % Ns)}).
// kFirst: the inputs come from the (windowed) audio instead of the ping-pong buffer.
// tws: this stage's twiddles, [r - 1][k] = e^{-2 pi i r k / (R Ns)}, k < Ns: lanes of a warp read consecutive k (no bank conflicts; a
// single table e^{-2 pi i m / M} indexed by r k M / (R Ns) puts all lanes of an early stage on one bank).
template <int R, bool kFirst>
__device__ __forceinline__ void stockham_stage(const float2 *__restrict__ in, float2 *__restrict__ out, int M, int Ns, int lane,
                                               const float2 *__restrict__ tws, const float *__restrict__ fr, const float *__restrict__ win, int W,
                                               bool pairs = false) {
    const int nb = M / R;            // butterflies of this stage
#ifndef WIDE_STAGE_UNROLL
#define WIDE_STAGE_UNROLL 1
#endif
    constexpr int kUnroll = WIDE_STAGE_UNROLL;   // 2 (both butterflies of a lane in flight, 127 registers) measured slower: 7.37 against 7.08 ms
#pragma unroll kUnroll
    for (int j = lane; j < nb; j += 32) {
        const int k = j & (Ns - 1);
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = j + r * nb;
            if constexpr (kFirst) {
                const int m0 = 2 * n;
                if (pairs && m0 + 1 < W) {   // even hop, even window: one 64-bit load each for the sample pair and its window taps
                    const float2 x2 = reinterpret_cast<const float2 *>(fr)[n], w2 = reinterpret_cast<const float2 *>(win)[n];
                    v[r] = make_float2(x2.x * w2.x, x2.y * w2.y);
                } else {
                    v[r] = make_float2(m0 < W ? fr[m0] * win[m0] : 0.0f, m0 + 1 < W ? fr[m0 + 1] * win[m0 + 1] : 0.0f);   // zero padding: CSTFT.swift:109-110
                }
            } else {
                v[r] = in[padf(n)];
            }
        }
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tws[(r - 1) * Ns + k]);
        }
        Dft<R>::run(v);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[padf(base + r * Ns)] = v[r];
    }
}

// The 512-point transform of BASELINE config 4 (FFT 1024, full-length window, even hop) with every address of the three radix-8
// stages written as (per-lane constant) ^ or + (immediate): the swizzle of padf() is linear over GF(2), so for element
// n = j + 64 r the position is (lane part) + 512 r bytes, XOR 96 for odd r (bit 6 of n), and the scattered stores of the first two
// stages are (lane part) ^ (8 r) and (lane part) ^ 8 (8 r ^ (r >> 1)). The generic stockham_stage spends more instructions on
// index arithmetic than on the butterflies (176 per 8-point butterfly in the second stage, 60 of them floating point).
// b0 / b1: this warp's ping-pong buffers (512-byte aligned); the transform ends in b1.
struct Fft512Lane {
    uint32_t ld[2], ld_odd[2];   // element j + 64 r, even / odd r, before the + 512 r
    uint32_t st1[2], st2[2];     // stores of stage 1 (index 8 j + r) and stage 2 (64 (j >> 3) + (j & 7) + 8 r) at r = 0
    uint32_t tw2, tw3[2];        // twiddle rows: (j & 7) and j, in bytes
    __device__ __forceinline__ void init(int lane) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const uint32_t j = lane + 32 * p;
            ld[p] = 8u * ((j & ~15u) | ((j & 15u) ^ ((j >> 4) & 3u)));
            ld_odd[p] = ld[p] ^ 96u;
            st1[p] = 8u * ((8u * j) ^ (((j >> 1) & 3u) ^ (((j >> 3) & 1u) * 12u)));
            st2[p] = 8u * ((64u * (j >> 3)) + ((j & 7u) ^ (((j >> 3) & 1u) * 12u)));
            tw3[p] = 8u * j;
        }
        tw2 = 8u * (lane & 7u);
    }
};
__device__ __forceinline__ float2 lds2(const unsigned char *base, uint32_t off) { return *reinterpret_cast<const float2 *>(base + off); }
__device__ __forceinline__ void sts2(unsigned char *base, uint32_t off, float2 v) { *reinterpret_cast<float2 *>(base + off) = v; }

__device__ __forceinline__ void fft512_frame(const Fft512Lane &L, const float *fr, const float *win, const float2 *twS, float2 *b0f, float2 *b1f) {
    unsigned char *b0 = reinterpret_cast<unsigned char *>(b0f), *b1 = reinterpret_cast<unsigned char *>(b1f);
    const unsigned char *tw = reinterpret_cast<const unsigned char *>(twS);
    const unsigned char *frb = reinterpret_cast<const unsigned char *>(fr), *winb = reinterpret_cast<const unsigned char *>(win);
    const int lane8 = (int)(L.tw3[0]);   // 8 * lane
    // stage 1: windowed sample pairs (x[2n], x[2n+1]) as complex inputs, n = j + 64 r; outputs at 8 j + r
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = ptx::mul2(lds2(frb, lane8 + 256 * p + 512 * r), lds2(winb, lane8 + 256 * p + 512 * r));
        Dft<8>::run(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) sts2(b1, L.st1[p] ^ (8u * r), v[r]);
    }
    __syncwarp();
    // stage 2 (Ns = 8): b1 -> b0, twiddles [r - 1][j & 7]
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = lds2(b1, ((r & 1) ? L.ld_odd[p] : L.ld[p]) + 512u * r);
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], lds2(tw, L.tw2 + 64u * (r - 1)));
        Dft<8>::run(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) sts2(b0, L.st2[p] ^ (8u * ((8u * r) ^ (uint32_t)(r >> 1))), v[r]);
    }
    __syncwarp();
    // stage 3 (Ns = 64): b0 -> b1, twiddles [r - 1][j] behind the 56 entries of stage 2; outputs at j + 64 r
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        float2 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = lds2(b0, ((r & 1) ? L.ld_odd[p] : L.ld[p]) + 512u * r);
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], lds2(tw, 56u * 8u + L.tw3[p] + 512u * (r - 1)));
        Dft<8>::run(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) sts2(b1, ((r & 1) ? L.ld_odd[p] : L.ld[p]) + 512u * r, v[r]);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(512) stft_planes_fast_kernel(const DevNet *__restrict__ netp, const float *__restrict__ pcm, int64_t ch_stride,
                                                               int64_t col0, int64_t n_cols, float *__restrict__ hi, float *__restrict__ lo,
                                                               float4 *__restrict__ stats, int n_planes, int64_t rows_alloc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevNet &net = *netp;
    const int N = net.fft_len, M = N / 2, L = net.band, W = net.win_len, hop = net.hop;
    const int warps = blockDim.x / 32, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int pitch = n_planes * 4 + 1;
    const int span = (kWideStftCols - 1) * hop + W;
    float *audio = reinterpret_cast<float *>(smem_raw);                  // [span]
    float *win = audio + ((span + 3) & ~3);                              // [W]
    float *tile = win + ((W + 3) & ~3);                                  // [kWideStftCols][pitch]
    float2 *twS = reinterpret_cast<float2 *>(tile + ((kWideStftCols * pitch + 1) & ~1));   // per-stage twiddle tables, < 2 M entries in all
    float2 *utw = twS + 2 * M;                                           // [L] e^{-2 pi i (k0 + f) / N}
    float2 *buf = utw + ((L + 1) & ~1);                                  // [warps][2][M], swizzled (padf), 512-byte aligned
    buf += ((512u - (ptx::smem_addr(buf) & 511u)) & 511u) / 8;
    const int bufM = M;
    float2 *b0 = buf + (size_t)warp * 2 * bufM, *b1 = b0 + bufM;
    const int ch = blockIdx.y;
    const int64_t c_first = (int64_t)blockIdx.x * kWideStftCols;
    const int cols = (int)min((int64_t)kWideStftCols, n_cols - c_first);
    const float *src = pcm + (int64_t)ch * ch_stride + (col0 + c_first) * hop + net.gap;
    const int need = (cols - 1) * hop + W;
    for (int i = threadIdx.x; i < need; i += blockDim.x) audio[i] = src[i];
    for (int i = threadIdx.x; i < W; i += blockDim.x) win[i] = net.window[i];
    for (int i = threadIdx.x; i < kWideStftCols * pitch; i += blockDim.x) tile[i] = 0.0f;   // padding bins and missing columns read as 0
    int log_m = 0;
    while ((1 << log_m) < M) ++log_m;
    const int last_radix = log_m % 3 == 0 ? 8 : (log_m % 3 == 1 ? 2 : 4);
    {   // twiddle tables of every stage after the first, back to back: stage with sub-transform length Ns and radix R at offset off
        int off = 0;
        for (int Ns = 8; Ns < M; Ns *= 8) {
            const int R = Ns * 8 <= M ? 8 : last_radix;
            for (int i = threadIdx.x; i < (R - 1) * Ns; i += blockDim.x) {
                const int r = i / Ns + 1, k = i % Ns;
                double sn, cs;
                sincospi(-2.0 * (double)(r * k) / (double)(R * Ns), &sn, &cs);
                twS[off + i] = make_float2((float)cs, (float)sn);
            }
            off += (R - 1) * Ns;
        }
    }
    for (int f = threadIdx.x; f < L; f += blockDim.x) {
        double sn, cs;
        sincospi(-2.0 * (double)(net.k0 + f) / (double)N, &sn, &cs);
        utw[f] = make_float2((float)cs, (float)sn);
    }
    __syncthreads();
    Fft512Lane l512;
    l512.init(lane);
    for (int c = warp; c < cols; c += warps) {
        const float *fr = audio + c * hop;
        float2 *in = b0, *out = b1;
        int Ns = 1;
        // radix-8 stages, the first one straight from the windowed audio; then the remainder stage
        if (M == 512 && W == 1024 && (hop & 1) == 0) {   // BASELINE config 4: hand-addressed transform
            fft512_frame(l512, fr, win, twS, b0, b1);
            out = b1;
        } else if (M == 512) {   // literal sizes, so the inlined stages fold their index arithmetic
            stockham_stage<8, true>(nullptr, b1, 512, 1, lane, twS, fr, win, W, (hop & 1) == 0);
            __syncwarp();
            stockham_stage<8, false>(b1, b0, 512, 8, lane, twS, nullptr, nullptr, 0);
            __syncwarp();
            stockham_stage<8, false>(b0, b1, 512, 64, lane, twS + 56, nullptr, nullptr, 0);
            __syncwarp();
            out = b1;
        } else if (M >= 8) {
            stockham_stage<8, true>(nullptr, out, M, 1, lane, twS, fr, win, W, (hop & 1) == 0);
            Ns = 8;
            int off = 0;
            __syncwarp();
            while (Ns * (last_radix == 8 ? 1 : last_radix) < M) {
                float2 *t2 = in; in = out; out = t2;
                stockham_stage<8, false>(in, out, M, Ns, lane, twS + off, nullptr, nullptr, 0);
                off += 7 * Ns;
                Ns *= 8;
                __syncwarp();
            }
            if (last_radix != 8) {
                float2 *t2 = in; in = out; out = t2;
                if (last_radix == 2) stockham_stage<2, false>(in, out, M, Ns, lane, twS + off, nullptr, nullptr, 0);
                else stockham_stage<4, false>(in, out, M, Ns, lane, twS + off, nullptr, nullptr, 0);
                __syncwarp();
            }
        }
        // untangle the packed real transform, magnitude, band slice (as kernels_fused.cu): X[k] = ((Z[k] + conj Z[M-k]) - i w^k (Z[k] - conj Z[M-k])) / 2;
        // with M-k taken mod M the same expression gives |Re Z[0] + Im Z[0]| for k = 0 (the Nyquist term is dropped upstream, CSTFT.swift:323)
        const float2 *z = out;
        float *row = tile + c * pitch;
        float ss = 0.0f, mn = INFINITY, mx = -INFINITY;
        if (net.scaling == SYLDET_SCALING_LINEAR) {
            // the common case without the scaling dispatch, with packed adds and the special-function square root (relative
            // error 2^-23: the band magnitudes stay within the 1e-5 spectrum tolerance of the oracle's correctly rounded sqrtf)
            for (int f = lane; f < L; f += 32) {
                const int k = net.k0 + f, km = (M - k) & (M - 1);
                const float2 za = z[padf(k)], zc = z[padf(km)], t = utw[f];
                const float2 zcc = make_float2(zc.x, -zc.y);                     // conj Z[M - k]
                const float2 sm = ptx::add2(za, zcc), df = ptx::sub2(za, zcc);   // (sr, si), (dr, di)
                const float re = sm.x + fmaf(t.x, df.y, t.y * df.x), im = sm.y - fmaf(t.x, df.x, -(t.y * df.y));
                const float v = 0.5f * sqrt_fast(fmaf(re, re, im * im));
                row[f] = v;
                ss = fmaf(v, v, ss);
                mn = fminf(mn, v);
                mx = fmaxf(mx, v);
            }
        } else
        for (int f = lane; f < L; f += 32) {
            const int k = net.k0 + f, km = (M - k) & (M - 1);
            const float2 za = z[padf(k)], zc = z[padf(km)], t = utw[f];
            const float sr = za.x + zc.x, si = za.y - zc.y, dr = za.x - zc.x, di = za.y + zc.y;
            const float re = sr + (t.x * di + t.y * dr), im = si - (t.x * dr - t.y * di);
            float v = 0.5f * sqrtf(re * re + im * im);
            v = scale_value(v, net.scaling);
            row[f] = v;
            ss = fmaf(v, v, ss);
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
        __syncwarp();
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
            lo4[(int64_t)pl * rows_alloc + r] = make_float4(raw.x - tf32_trunc_w(raw.x), raw.y - tf32_trunc_w(raw.y), raw.z - tf32_trunc_w(raw.z),
                                                            raw.w - tf32_trunc_w(raw.w));
        }
    }
}

}  // namespace

size_t wide_smem_bytes(int h_pad) { return 128 + WideSmem::total(h_pad); }
int wide_tile_rows() { return kWTileRows; }
int wide_max_time_range() { return kWRowsPad - kWTileRows + 1; }
size_t wide_weight_block_bytes() { return kWWBytes; }

bool stft_planes_fast_supported(int fft_len) { return fft_len >= 16 && fft_len <= 2048; }

static size_t stft_planes_fast_smem(int fft_len, int win_len, int band, int hop, int n_planes, int warps) {
    const int M = fft_len / 2, span = (kWideStftCols - 1) * hop + win_len, pitch = n_planes * 4 + 1;
    const size_t floats = (size_t)((span + 3) & ~3) + ((win_len + 3) & ~3) + ((kWideStftCols * pitch + 1) & ~1);
    const size_t f2 = (size_t)2 * M + ((band + 1) & ~1) + (size_t)warps * 2 * M;
    return floats * 4 + f2 * 8 + 16 + 512;   // + alignment of the ping-pong buffers
}

cudaError_t launch_stft_planes_fast(const DevNet *d_net, int fft_len, int win_len, int band, int hop, const float *pcm, int64_t ch_stride,
                                    int n_channels, int64_t col0, int64_t n_cols, float *hi, float *lo, float4 *stats, int n_planes,
                                    int64_t rows_alloc, cudaStream_t stream) {
    if (n_cols <= 0 || n_channels <= 0) return cudaSuccess;
    int warps = 16;   // one CTA per SM: as many warps as the ping-pong buffers leave room for
    while (warps > 4 && stft_planes_fast_smem(fft_len, win_len, band, hop, n_planes, warps) > 220 * 1024) warps -= 4;
    const size_t smem = stft_planes_fast_smem(fft_len, win_len, band, hop, n_planes, warps);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(stft_planes_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((n_cols + kWideStftCols - 1) / kWideStftCols), (unsigned)n_channels);
    stft_planes_fast_kernel<<<grid, warps * 32, smem, stream>>>(d_net, pcm, ch_stride, col0, n_cols, hi, lo, stats, n_planes, rows_alloc);
    return cudaGetLastError();
}

cudaError_t launch_wide(int grid, const WideParams &p, const WideWork &w, cudaStream_t stream) {
    const size_t smem = wide_smem_bytes(p.h_pad);
    cudaError_t e = cudaFuncSetAttribute(wide_l0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    wide_l0_kernel<<<grid, kWThreads, smem, stream>>>(p, w);
    return cudaGetLastError();
}

}  // namespace syldet
