// Fused sm_100a kernel: audio tile -> Hamming window -> real FFT -> band magnitude -> sliding feature window ->
// input normalisation -> small MLP -> threshold decision, with spectra and features kept in shared memory.
//
// Replaces, for all channels and all hops of a recording at once, the per-column loop
//   CircularShortTimeFourierTransform.extractPower   (Common/CircularShortTimeFourierTransform.swift:280-337)
//   SyllableDetector.processFourierData/processNewValue (Common/SyllableDetector.swift:134-217)
//   NeuralNet.apply / NeuralNetLayer.apply             (Common/NeuralNet.swift:294-326, 366-377)
//   the threshold test of TrackDetector.process        (SyllableDetectorCLI/TrackDetector.swift:71-77)
//
// Work decomposition (B200: 148 SMs, 2 CTAs of 256 threads per SM):
//   unit   = (channel, chunk of consecutive evaluations); CTAs walk units grid-stride.
//   round  = 8 warps x G frames; the audio span of a round is staged in shared memory by a 1-D TMA bulk copy
//            (cp.async.bulk + mbarrier), double buffered against the FFT of the previous round.
//   FFT    = real N-point transform as an N/2-point complex Stockham transform in two register passes
//            (radix R1 then R2, one padded shared-memory exchange, warp-private, __syncwarp only).
//   ring   = band magnitudes of the last nn_tile + T columns (shared memory, odd pitch).
//   epilogue = one thread per evaluation: gathers its T columns from the ring, layer-0 dot products against folded
//            weights that sit in kernel-parameter constant memory (warp-uniform FFMA operands), window statistic,
//            transfer functions, remaining layers, reverse output map, double-precision threshold compare,
//            warp-aggregated event append.
#include <cstdio>

#include "fft_regs.cuh"
#include "fused_epilogue.cuh"
#include "ptx_sm100.cuh"

namespace syldet {

namespace {

// Where the audio of one round sits in global memory and in the staging buffer.
struct RoundSpan {
    const float *src;     // first sample of the round's first frame
    const char *aligned;  // 16-byte aligned start of the bulk copy
    int off;              // floats between `aligned` and `src` (0..3)
    uint32_t bytes;       // bulk copy size (multiple of 16)
    int n_floats;         // samples the round's frames touch
    bool bulk_ok;         // the aligned span lies inside the caller's buffer
};

__device__ __forceinline__ RoundSpan round_span(const FusedParams &p, const FusedWork &w, const float *src, int cols) {
    RoundSpan s;
    s.src = src;
    s.n_floats = (cols - 1) * p.hop + p.win_len;
    const uintptr_t a = (uintptr_t)src;
    s.aligned = (const char *)(a & ~(uintptr_t)15);
    s.off = (int)((a & 15) >> 2);
    s.bytes = (uint32_t)(((s.off + s.n_floats) * 4 + 15) & ~15);
    s.bulk_ok = (uintptr_t)s.aligned >= (uintptr_t)w.pcm_begin && (uintptr_t)s.aligned + s.bytes <= (uintptr_t)w.pcm_end;
    return s;
}

// ---- the kernel --------------------------------------------------------------------------------------------------
template <int NFFT, int HP>
__global__ void __launch_bounds__(kFusedThreads, 2) fused_detect_kernel(const __grid_constant__ FusedParams p, const FusedWork w) {
    constexpr int M = NFFT / 2;
    constexpr int R1 = fused_r1(NFFT), R2 = fused_r2(NFFT);
    constexpr int G = fused_group(NFFT);             // frames per warp per round
    constexpr int RC = fused_round_cols(NFFT);       // columns per round (CTA)
    constexpr int FP = fused_frame_pitch(NFFT);      // float2 per frame in the exchange buffer
    constexpr int U2 = (G * R1) / 32;                // pass-2 items per lane
    static_assert(R1 * R2 == M && (G * M) / R1 == 32 && U2 >= 1, "plan");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);
    float2 *s_utw = reinterpret_cast<float2 *>(smem_raw + 16);             // untangle twiddles w_N^{k0+f}, f < band
    float *abuf0 = reinterpret_cast<float *>(smem_raw + 16 + kFusedMaxBand * sizeof(float2));
    float *abuf1 = abuf0 + p.abuf_floats;
    float2 *scratch = reinterpret_cast<float2 *>(abuf1 + p.abuf_floats);
    float *ring = reinterpret_cast<float *>(scratch + (kFusedThreads / 32) * G * FP);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.band, T = p.time_range, W = p.win_len;
    const bool full_window = (W == NFFT);
    constexpr int LOG_R1 = R1 == 16 ? 4 : 3;

    // per-lane constants -----------------------------------------------------------------------------------------
    const int fs1 = lane / R2, j1 = lane % R2;  // pass 1: frame in group, butterfly index (M/R1 == R2 items per frame)
    float2 wreg[R1];                            // window taps of this lane's samples
#pragma unroll
    for (int r = 0; r < R1; ++r) {
        const int m0 = 2 * (j1 + r * R2);
        wreg[r].x = m0 < W ? __ldg(w.window + m0) : 0.0f;
        wreg[r].y = m0 + 1 < W ? __ldg(w.window + m0 + 1) : 0.0f;
    }
    const int j2 = lane % R1;                   // pass 2
    float2 tw[R2];                              // w_M^{r j2} = w_N^{2 r j2}
#pragma unroll
    for (int r = 1; r < R2; ++r) {
        const int q = 2 * r * j2;  // < N; the table holds k < N/2 and w_N^{k + N/2} = -w_N^k
        const float2 t = __ldg(w.twiddle + (q >= M ? q - M : q));
        tw[r] = q >= M ? make_float2(-t.x, -t.y) : t;
    }
    for (int f = tid; f < L; f += kFusedThreads) s_utw[f] = __ldg(w.twiddle + p.k0 + f);

    if (tid == 0) {
        ptx::mbar_init(&mbar[0], 1);
        ptx::mbar_init(&mbar[1], 1);
        ptx::fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;

    float2 *zw = scratch + warp * G * FP;  // this warp's exchange buffer
    const int64_t n_units = (int64_t)w.n_channels * w.chunks_per_channel;
    const int round_floats = RC * p.hop;   // samples between the first frames of consecutive rounds
    const int nq = (L + 31) >> 5;          // band slots per lane

    for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int ch = (int)(unit / w.chunks_per_channel);
        const int64_t e0 = (unit - (int64_t)ch * w.chunks_per_channel) * w.chunk_evals;
        const int ne = (int)min(w.chunk_evals, w.evals_per_channel - e0);
        const int ncols = ne + T - 1;  // column index == evaluation index of the window it opens
        const int nrounds = (ncols + RC - 1) / RC;
        const float *src = w.pcm + (int64_t)ch * w.ch_stride + (w.eval_offset + e0) * p.hop + p.gap;  // first sample of this round's first frame
        float *out_base = w.all_out ? w.all_out + ((int64_t)ch * w.out_evals_per_channel + w.eval_offset + e0) * p.n_out : nullptr;

        if (tid == 0) {
            const RoundSpan s = round_span(p, w, src, min(RC, ncols));
            if (s.bulk_ok) {
                ptx::mbar_expect_tx(&mbar[0], s.bytes);
                ptx::bulk_copy_g2s(abuf0, s.aligned, s.bytes, &mbar[0]);
            }
        }
        int cols_done = 0, evals_done = 0;
        int col_slot = 0, eval_slot = 0;  // ring slots of column `cols_done` and evaluation `evals_done`

        for (int r = 0; r < nrounds; ++r, src += round_floats) {
            const int cols = min(RC, ncols - r * RC);
            float *abuf = (r & 1) ? abuf1 : abuf0;
            if (tid == 0 && r + 1 < nrounds) {  // prefetch next round into the other buffer (free since the last sync)
                const RoundSpan s = round_span(p, w, src + round_floats, min(RC, ncols - (r + 1) * RC));
                if (s.bulk_ok) {
                    uint64_t *bar = &mbar[(r + 1) & 1];
                    ptx::mbar_expect_tx(bar, s.bytes);
                    ptx::bulk_copy_g2s((r & 1) ? abuf0 : abuf1, s.aligned, s.bytes, bar);
                }
            }
            const RoundSpan span = round_span(p, w, src, cols);
            if (span.bulk_ok) {
                if (r & 1) { ptx::mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
                else { ptx::mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
            } else {  // edge of the caller's buffer: guarded cooperative copy
                for (int i = tid; i < span.n_floats; i += kFusedThreads) abuf[span.off + i] = span.src[i];
                __syncthreads();
                ptx::fence_proxy_async_smem();
            }

            // ---- pass 1: window + radix-R1 on strided samples -------------------------------------------------
            {
                const int c = warp * G + fs1;
                if (c < cols) {
                    const int first = span.off + c * p.hop;
                    const float *fr = abuf + first + 2 * j1;
                    float2 v[R1];
                    if (full_window && (first & 1) == 0) {  // 8-byte aligned pairs: one LDS.64 per complex point
#pragma unroll
                        for (int r1 = 0; r1 < R1; ++r1) {
                            const float2 x = *reinterpret_cast<const float2 *>(fr + 2 * r1 * R2);
                            v[r1] = make_float2(x.x * wreg[r1].x, x.y * wreg[r1].y);
                        }
                    } else if (full_window) {
#pragma unroll
                        for (int r1 = 0; r1 < R1; ++r1)
                            v[r1] = make_float2(fr[2 * r1 * R2] * wreg[r1].x, fr[2 * r1 * R2 + 1] * wreg[r1].y);
                    } else {  // window shorter than the FFT: zero padding (CSTFT.swift:109-110)
#pragma unroll
                        for (int r1 = 0; r1 < R1; ++r1) {
                            const int m0 = 2 * (j1 + r1 * R2);
                            const float x0 = m0 < W ? fr[2 * r1 * R2] : 0.0f;
                            const float x1 = m0 + 1 < W ? fr[2 * r1 * R2 + 1] : 0.0f;
                            v[r1] = make_float2(x0 * wreg[r1].x, x1 * wreg[r1].y);
                        }
                    }
                    Dft<R1>::run(v);
                    float2 *zb = zw + fs1 * FP + j1 * (R1 + 1);  // padded: element e lives at e + e / R1
#pragma unroll
                    for (int q = 0; q < R1; ++q) zb[q] = v[q];
                }
            }
            __syncwarp();
            // ---- pass 2: twiddle + radix-R2, in place ---------------------------------------------------------
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int fs2 = lane / R1 + u * (32 / R1);
                if (warp * G + fs2 < cols) {
                    float2 *zb = zw + fs2 * FP + j2;
                    float2 v[R2];
                    v[0] = zb[0];
#pragma unroll
                    for (int r2 = 1; r2 < R2; ++r2) v[r2] = cmul(zb[r2 * (R1 + 1)], tw[r2]);
                    Dft<R2>::run(v);
#pragma unroll
                    for (int r2 = 0; r2 < R2; ++r2) zb[r2 * (R1 + 1)] = v[r2];
                }
            }
            __syncwarp();
            // ---- untangle the packed real transform, magnitude, band slice -> ring -------------------------------
            // X[k] = ((Z[k] + conj Z[M-k]) - i w^k (Z[k] - conj Z[M-k])) / 2; with M-k taken mod M the same expression
            // gives |Re Z[0] + Im Z[0]| for k = 0 (the Nyquist term is dropped upstream, CSTFT.swift:323).
            {
                const int gmax = min(G, cols - warp * G);
                int slot = col_slot + warp * G;
                if (slot >= p.ring_cols) slot -= p.ring_cols;
                for (int g = 0; g < gmax; ++g) {
                    const float2 *zb = zw + g * FP;
                    float *dst = ring + slot * p.band_pitch;
                    for (int q = 0; q < nq; ++q) {
                        const int f = lane + 32 * q;
                        const int fc = min(f, L - 1);
                        const int k = p.k0 + fc, km = (M - k) & (M - 1);
                        const float2 za = zb[k + (k >> LOG_R1)];
                        const float2 zc = zb[km + (km >> LOG_R1)];
                        const float2 t = s_utw[fc];
                        const float sr = za.x + zc.x, si = za.y - zc.y;  // Z[k] + conj Z[M-k]
                        const float dr = za.x - zc.x, di = za.y + zc.y;  // Z[k] - conj Z[M-k]
                        const float re = sr + (t.x * di + t.y * dr);
                        const float im = si - (t.x * dr - t.y * di);
                        float mag = 0.5f * sqrt_fast(re * re + im * im);
                        if (w.debug_band && f < L)   // extractPower() values, before the scaling (column index == evaluation index)
                            w.debug_band[((int64_t)ch * w.debug_cols + w.eval_offset + e0 + r * RC + warp * G + g) * L + f] = mag;
                        if (p.scaling != SYLDET_SCALING_LINEAR) mag = scale_value(mag, p.scaling);
                        if (f < L) dst[f] = mag;
                    }
                    if (++slot == p.ring_cols) slot = 0;
                }
            }
            __syncthreads();
            cols_done += cols;
            col_slot += cols;
            if (col_slot >= p.ring_cols) col_slot -= p.ring_cols;

            // ---- epilogue over the evaluations whose T columns are complete ---------------------------------------
            const int n_ready = cols_done - (T - 1) - evals_done;
            if (n_ready >= p.nn_tile || (r == nrounds - 1 && n_ready > 0)) {
                for (int qb = warp * 32; qb < n_ready; qb += kFusedThreads) {
                    const int q = qb + lane;
                    const bool active = q < n_ready;
                    float out[kFusedMaxOut];
                    bool hit = false;
                    if (active) {
                        int slot = eval_slot + q;
                        if (slot >= p.ring_cols) slot -= p.ring_cols;
                        hit = evaluate<HP>(p, w.detect_rule, ring, slot, out);
                        if (out_base) {
                            float *o = out_base + (int64_t)(evals_done + q) * p.n_out;
#pragma unroll
                            for (int i = 0; i < kFusedMaxOut; ++i)
                                if (i < p.n_out) o[i] = out[i];
                        }
                    }
                    const unsigned hits = __ballot_sync(0xffffffffu, hit);
                    if (hits) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(w.sink.count, (unsigned long long)__popc(hits));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (hit) {
                            const unsigned long long idx = base + __popc(hits & ((1u << lane) - 1));
                            if (idx < w.sink.capacity) {
                                w.sink.events[idx] = DevEvent{ch, 0, w.eval_offset + e0 + evals_done + q};
#pragma unroll
                                for (int i = 0; i < kFusedMaxOut; ++i)
                                    if (i < p.n_out) w.sink.outputs[idx * p.n_out + i] = out[i];
                            }
                        }
                    }
                }
                evals_done += n_ready;
                eval_slot += n_ready;
                while (eval_slot >= p.ring_cols) eval_slot -= p.ring_cols;
            }
        }
        __syncthreads();  // ring and staging buffers are reused by the next unit
    }
}

template <int NFFT, int HP>
cudaError_t launch_one(const FusedLaunch &cfg, const FusedParams &p, const FusedWork &w, cudaStream_t stream) {
    auto kern = fused_detect_kernel<NFFT, HP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
    if (e != cudaSuccess) return e;
    kern<<<cfg.grid, kFusedThreads, cfg.smem, stream>>>(p, w);
    return cudaGetLastError();
}

template <int NFFT, int HP>
cudaError_t occupancy_one(size_t smem, int *blocks) {
    auto kern = fused_detect_kernel<NFFT, HP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kern, kFusedThreads, smem);
}

}  // namespace

bool fused_supports_fft(int fft_len) { return fft_len == 64 || fft_len == 128 || fft_len == 256 || fft_len == 512; }

size_t fused_smem_bytes(int fft_len, const FusedParams &p) {
    const size_t scratch = (size_t)(kFusedThreads / 32) * fused_group(fft_len) * fused_frame_pitch(fft_len) * sizeof(float2);
    return 16 + kFusedMaxBand * sizeof(float2) + 2 * (size_t)p.abuf_floats * sizeof(float) + scratch + (size_t)p.ring_cols * p.band_pitch * sizeof(float);
}

#define SYLDET_FUSED_DISPATCH(FN, ...)                                             \
    switch (cfg_fft * 16 + cfg_hp) {                                               \
        case 64 * 16 + 4: return FN<64, 4>(__VA_ARGS__);                           \
        case 64 * 16 + 8: return FN<64, 8>(__VA_ARGS__);                           \
        case 128 * 16 + 4: return FN<128, 4>(__VA_ARGS__);                         \
        case 128 * 16 + 8: return FN<128, 8>(__VA_ARGS__);                         \
        case 256 * 16 + 4: return FN<256, 4>(__VA_ARGS__);                         \
        case 256 * 16 + 8: return FN<256, 8>(__VA_ARGS__);                         \
        case 512 * 16 + 4: return FN<512, 4>(__VA_ARGS__);                         \
        case 512 * 16 + 8: return FN<512, 8>(__VA_ARGS__);                         \
        default: return cudaErrorInvalidValue;                                     \
    }

cudaError_t launch_fused(const FusedLaunch &cfg, const FusedParams &p, const FusedWork &w, cudaStream_t stream) {
    const int cfg_fft = cfg.fft_len, cfg_hp = cfg.hp;
    SYLDET_FUSED_DISPATCH(launch_one, cfg, p, w, stream)
}

cudaError_t fused_max_blocks_per_sm(int fft_len, int hp, size_t smem, int *blocks) {
    const int cfg_fft = fft_len, cfg_hp = hp;
    SYLDET_FUSED_DISPATCH(occupancy_one, smem, blocks)
}

}  // namespace syldet
