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
#include <emmintrin.h>
#include <cstring>
#include <cstddef>
#include <cstdio>

#include "fft_regs.cuh"
#include "fused_epilogue.cuh"
#include "ptx_sm100.cuh"
#include "resample.cuh"

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

// ---- live tick, latency-shaped variant ------------------------------------------------------------------------------------
// stream_tick_kernel (kernels_generic.cu) does one tick of the live path in the reference's operation order (iterative radix-2 FFT in
// shared memory, the literal processing chain). Configurations the fused kernel takes get this kernel instead - same protocol
// (samples pulled from pinned host staging into the device ring, band columns into the device band ring, outputs + sequence number to
// pinned host memory, level meters, resampler) - built for LATENCY. A tick is one column and one evaluation on a single block per
// channel, so nothing hides behind other warps: the time is (PCIe round trip of the staged samples) + (instructions on the longest
// dependent path) x ~3 cycles. Hence:
//   * every load the tick needs is issued at the top, behind the PCIe reads of the staged samples: the older samples of the frames
//     and the older band columns of the evaluation windows go ring -> shared memory with cp.async (no registers, no scoreboard), the
//     window taps, twiddles and this thread's share of the folded layer-0 weights go to registers;
//   * what the tick itself produces (new samples, new columns) is handed on in shared memory - never store -> L2 -> load;
//   * frames and evaluation windows are CONTIGUOUS in shared memory (`span`: samples from the first frame's start; `win`: columns
//     from the first window's start), so the hot loops are immediate-offset LDS + packed arithmetic with no index bookkeeping;
//   * the folded layer 0 of an evaluation is split over all 128 threads (each holds its weights from the top of the kernel), then
//     one warp finishes the network.
constexpr int kTickThreads = 128;
constexpr int kTickPre = 4;        // staged samples per thread held in registers across the PCIe round trip (512 per channel): one float4
constexpr int kTickEvalGroup = 4;  // evaluations reduced per block-wide pass

// Per-thread invariants of a configuration: this thread's share of the folded layer-0 weights, and for the warps that transform
// frames the window taps, pass-2 twiddles and untangle twiddles. A launched tick loads them behind its PCIe reads; the resident
// kernel loads them once.
template <int NFFT, int HP>
struct TickInv {
    static constexpr int kW = kFusedMaxW0 / HP / kTickThreads;   // layer-0 inputs per thread (12 or 6): the whole layer in one pass
    static constexpr int kBins = (kFusedMaxBand + 31) / 32;
    float4 wa[kW], wb[HP == 8 ? kW : 1];
    float2 wreg[fused_r1(NFFT)], tw[fused_r2(NFFT)], utw[kBins];
};

template <int NFFT, int HP>
__device__ __forceinline__ void tick_load_inv(TickInv<NFFT, HP> &r, const FusedParams *__restrict__ pp, const TickGeom &g, const float *__restrict__ window,
                                              const float2 *__restrict__ twiddle, bool want_weights, bool want_cols) {
    constexpr int M = NFFT / 2, R1 = fused_r1(NFFT), R2 = fused_r2(NFFT);
    const int tid = threadIdx.x, lane = tid & 31;
    const int W = g.win_len, L = g.band, I = g.band * g.time_range;
    if (want_weights) {
#pragma unroll
        for (int k = 0; k < TickInv<NFFT, HP>::kW; ++k) {
            const int i = tid + k * kTickThreads;
            if (i < I) {
                r.wa[k] = __ldg(reinterpret_cast<const float4 *>(pp->w0 + (size_t)i * HP));
                if constexpr (HP == 8) r.wb[k] = __ldg(reinterpret_cast<const float4 *>(pp->w0 + (size_t)i * HP + 4));
            }
        }
    }
    if (want_cols) {
        const int j1 = lane % R2, j2 = lane % R1;
#pragma unroll
        for (int q = 0; q < R1; ++q) {
            const int m0 = 2 * (j1 + q * R2);
            r.wreg[q].x = m0 < W ? __ldg(window + m0) : 0.0f;       // zero padding: CSTFT.swift:109-110
            r.wreg[q].y = m0 + 1 < W ? __ldg(window + m0 + 1) : 0.0f;
        }
#pragma unroll
        for (int q = 1; q < R2; ++q) {
            const int e = 2 * q * j2;
            const float2 v = __ldg(twiddle + (e >= M ? e - M : e));
            r.tw[q] = e >= M ? make_float2(-v.x, -v.y) : v;
        }
        r.tw[0] = make_float2(1.0f, 0.0f);
#pragma unroll
        for (int q = 0; q < TickInv<NFFT, HP>::kBins; ++q) r.utw[q] = lane + 32 * q < L ? __ldg(twiddle + g.k0 + lane + 32 * q) : make_float2(0.f, 0.f);
    }
}

struct TickSmem {
    FusedParams *sp;   // only the head (everything in front of w0) exists here
    float *span, *win, *stage, *part;
    float2 *zw;        // [4 warps][G][frame pitch] exchange buffers
};
template <int NFFT, int HP>
__device__ __forceinline__ TickSmem tick_smem_layout(unsigned char *base, int head_bytes, const TickGeom &g) {
    TickSmem m;
    m.sp = reinterpret_cast<FusedParams *>(base);
    m.span = reinterpret_cast<float *>(base + head_bytes);
    m.win = m.span + g.span_floats;
    m.stage = m.win + g.win_floats;
    m.part = m.stage + g.stage_floats;
    m.zw = reinterpret_cast<float2 *>(m.part + kTickEvalGroup * 4 * (HP + 2));
    return m;
}

__device__ __forceinline__ float ld_volatile_f32(const float *p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// One tick of one channel by one block. kResident: the block outlives the tick (stream_tick_resident_kernel): the invariants and the
// head of the parameter block are already in place, and pinned host memory is read with volatile loads (L1 is not flushed between
// ticks).
template <int NFFT, int HP, bool kResident>
__device__ __forceinline__ void tick_run(const FusedParams *__restrict__ pp, const StreamTick &t, const TickGeom &g, TickInv<NFFT, HP> &inv,
                                         const TickSmem &sm, const float *__restrict__ window, const float2 *__restrict__ twiddle) {
    constexpr int M = NFFT / 2;
    constexpr int R1 = fused_r1(NFFT), R2 = fused_r2(NFFT);
    constexpr int G = fused_group(NFFT);
    constexpr int FP = fused_frame_pitch(NFFT);
    constexpr int U2 = (G * R1) / 32;
    constexpr int LOG_R1 = R1 == 16 ? 4 : 3;
    constexpr int kHead = ((int)offsetof(FusedParams, w0) + 15) & ~15;   // everything but the layer-0 weights: staged in shared memory
    constexpr int kW = TickInv<NFFT, HP>::kW;
    constexpr int kBins = TickInv<NFFT, HP>::kBins;
    constexpr int kPart = HP + 2;                                        // partial sums per warp and evaluation: HP dot products + 2 statistics
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ch = blockIdx.x;
    const bool stamp = t.stamps != nullptr && ch == 0 && tid == 0;   // SYLDET_STREAM_TIMING
    long long ts[5] = {0, 0, 0, 0, 0}, sub[6] = {0, 0, 0, 0, 0, 0};
    if (stamp) ts[0] = clock64();

    // (0) the longest-latency loads first: this channel's staged samples, straight out of pinned host memory
    const float *src = t.staged + (int64_t)ch * t.stage_pitch;
    float pre[kTickPre];   // samples 4 tid .. 4 tid + 3: one 16-byte read per thread (the staging rows are 128-byte aligned)
    {
        const int i = kTickPre * tid;
        if (i + kTickPre <= t.n_staged) {
            float4 v;
            if constexpr (kResident) asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i) : "memory");
            else v = *reinterpret_cast<const float4 *>(src + i);
            pre[0] = v.x; pre[1] = v.y; pre[2] = v.z; pre[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < kTickPre; ++k) pre[k] = i + k < t.n_staged ? (kResident ? ld_volatile_f32(src + i + k) : src[i + k]) : 0.0f;
        }
    }

    FusedParams *sp = sm.sp;
    float *s_span = sm.span, *s_win = sm.win, *s_stage = sm.stage, *s_part = sm.part;   // s_part: [kTickEvalGroup][4 warps][kPart]
    float2 *zw = sm.zw + warp * G * FP;                                                  // this warp's exchange buffer
    const int W = g.win_len, L = g.band, I = g.band * g.time_range;
    float *ring = t.ring + (int64_t)ch * (t.ring_mask + 1);
    float *band = t.band + (int64_t)ch * (t.band_mask + 1) * L;

    // everything already on the device: ring -> shared memory, asynchronously
    const int64_t span0 = t.col0 * g.hop + g.gap;   // absolute sample index of s_span[0]
    {
        const int64_t n_old64 = t.ring_pos - span0;
        const int n_old = n_old64 < 0 ? 0 : n_old64 > g.span_floats ? g.span_floats : (int)n_old64;
        for (int i = tid; i < n_old; i += kTickThreads) ptx::cp_async4(s_span + i, ring + ((span0 + i) & t.ring_mask));
        if (t.n_evals > 0) {   // columns [eval0, col0) of the band ring are consecutive rows, wrapping once at most
            const int total = (int)(t.col0 - t.eval0) * L, ringf = (int)(t.band_mask + 1) * L;
            const int row0 = (int)(t.eval0 & t.band_mask) * L;
            for (int i = tid; i < total; i += kTickThreads) {
                int s = row0 + i;
                if (s >= ringf) s -= ringf;
                ptx::cp_async4(s_win + i, band + s);
            }
        }
        if constexpr (!kResident)
            for (int i = tid; i < kHead / 16; i += kTickThreads)
                ptx::cp_async16(reinterpret_cast<unsigned char *>(sp) + 16 * i, reinterpret_cast<const int4 *>(pp) + i);
    }
    // registers: this thread's share of the folded layer-0 weights; window taps and twiddles for the warps that transform a frame
    const int fs1 = lane / R2, j1 = lane % R2, j2 = lane % R1;
    if constexpr (!kResident) tick_load_inv<NFFT, HP>(inv, pp, g, window, twiddle, t.n_evals > 0, (int64_t)warp * G < t.n_cols);
    const float4 *wa = inv.wa, *wb = inv.wb;
    const float2 *wreg = inv.wreg, *tw = inv.tw, *utw = inv.utw;
    const float rs_last = t.rs_on ? __ldcg(t.rs_last_in + ch) : 0.0f;   // written by the previous tick (of this very block when resident)
    if (stamp) ts[1] = clock64();

    // (1) staged samples: into shared memory, and (no resampler) into the device ring and the frame span
    auto place = [&](int64_t pos, float v) {   // one sample at ring rate
        ring[pos & t.ring_mask] = v;
        const int64_t rel = pos - span0;
        if (rel >= 0 && rel < g.span_floats) s_span[rel] = v;
    };
#pragma unroll
    for (int k = 0; k < kTickPre; ++k) {
        const int i = kTickPre * tid + k;
        if (i < t.n_staged) {
            s_stage[i] = pre[k];
            if (!t.rs_on) place(t.ring_pos + i, pre[k]);
        }
    }
    for (int i = tid + kTickPre * kTickThreads; i < t.n_staged; i += kTickThreads) {
        const float v = kResident ? ld_volatile_f32(src + i) : src[i];
        s_stage[i] = v;
        if (!t.rs_on) place(t.ring_pos + i, v);
    }
    ptx::cp_async_wait_all();
    __syncthreads();
    const FusedParams &p = *sp;
    if (t.rs_on) {
        // device-rate buffers -> ResamplerLinear -> sample ring: one resampleVector call per staged buffer (Processor.swift:116-121)
        for (int b = 0; b < t.n_marks; ++b) {
            const int lo = b ? t.marks[b - 1] : 0, n_in = t.marks[b] - lo, n_out = t.rs_n_out[b];
            const float off = t.rs_offset[b];
            const float last = b ? s_stage[lo - 1] : rs_last;
            for (int k = tid; k < n_out; k += kTickThreads)
                place(t.ring_pos + t.rs_out0[b] + k, resample_linear_point(s_stage + lo, n_in, off, t.rs_step, last, off < 0.0f, k));
        }
        if (tid == 0 && t.n_staged > 0) t.rs_last_out[ch] = s_stage[t.n_staged - 1];
        __syncthreads();
    }
    if (stamp) ts[2] = clock64();
    // input level meter (Processor.swift:110-113, StatMax of the buffer's mean square; the meter sees the device-rate samples): one warp
    // per staged buffer, from the back - warp 0 goes straight to the first frame
    if (t.level_in && warp > 0) {
        for (int b = 0; b < t.n_marks; ++b) {
            if (3 - (b % 3) != warp) continue;
            const int lo = b ? t.marks[b - 1] : 0, hi = t.marks[b];
            float part = 0.0f;
            for (int i = lo + lane; i < hi; i += 32) part = fmaf(s_stage[i], s_stage[i], part);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);   // vDSP_svesq: summation order unspecified
            const double ms = (double)part / (double)(hi - lo);
            if (lane == 0 && hi > lo && ms == ms) atomicMax(t.level_in + ch, (unsigned long long)__double_as_longlong(ms));
        }
    }
    if (stamp) sub[0] = clock64();

    // (2) STFT columns: window + real N-point FFT as an N/2-point complex transform in two register passes (as fused_detect_kernel)
    const int64_t win_col0 = t.col0 - t.eval0;   // column of s_win the first new column lands in
    for (int64_t cb = (int64_t)warp * G; cb < t.n_cols; cb += (int64_t)4 * G) {
        if (cb + fs1 < t.n_cols) {
            const float *xs = s_span + (cb + fs1) * g.hop + 2 * j1;
            float2 v[R1];
            if (W == NFFT) {
#pragma unroll
                for (int r1 = 0; r1 < R1; ++r1) v[r1] = ptx::mul2(make_float2(xs[2 * r1 * R2], xs[2 * r1 * R2 + 1]), wreg[r1]);
            } else {
#pragma unroll
                for (int r1 = 0; r1 < R1; ++r1) {
                    const int m0 = 2 * (j1 + r1 * R2);
                    v[r1] = ptx::mul2(make_float2(m0 < W ? xs[2 * r1 * R2] : 0.0f, m0 + 1 < W ? xs[2 * r1 * R2 + 1] : 0.0f), wreg[r1]);
                }
            }
            Dft<R1>::run(v);
            float2 *zb = zw + fs1 * FP + j1 * (R1 + 1);
#pragma unroll
            for (int q = 0; q < R1; ++q) zb[q] = v[q];
        }
        __syncwarp();
        if (stamp) sub[1] = clock64();
#pragma unroll
        for (int u = 0; u < U2; ++u) {
            const int fs2 = lane / R1 + u * (32 / R1);
            if (cb + fs2 < t.n_cols) {
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
        if (stamp) sub[2] = clock64();
        const int gmax = (int)min((int64_t)G, t.n_cols - cb);
        for (int gi = 0; gi < gmax; ++gi) {
            const float2 *zb = zw + gi * FP;
            float *dst = band + ((t.col0 + cb + gi) & t.band_mask) * L;
            float *dwin = s_win + (win_col0 + cb + gi) * L;
            const bool to_win = t.n_evals > 0 && (win_col0 + cb + gi + 1) * L <= g.win_floats;
#pragma unroll
            for (int q = 0; q < kBins; ++q) {
                const int f = lane + 32 * q;
                if (f < L) {
                    const int k = g.k0 + f, km = (M - k) & (M - 1);
                    const float2 za = zb[k + (k >> LOG_R1)], zc = zb[km + (km >> LOG_R1)], tk = utw[q];
                    const float sr = za.x + zc.x, si = za.y - zc.y, dr = za.x - zc.x, di = za.y + zc.y;
                    const float re = sr + (tk.x * di + tk.y * dr), im = si - (tk.x * dr - tk.y * di);
                    float mag = 0.5f * sqrt_fast(re * re + im * im);
                    if (p.scaling != SYLDET_SCALING_LINEAR) mag = scale_value(mag, p.scaling);
                    dst[f] = mag;
                    if (to_win) dwin[f] = mag;
                }
            }
        }
        __syncwarp();
        if (stamp) sub[3] = clock64();
    }
    __syncthreads();
    if (stamp) ts[3] = clock64();

    // (3) evaluations: the folded layer 0 of each split over the whole block, then one warp per evaluation finishes the network
    const int O = p.n_out;
    for (int64_t j0 = 0; j0 < t.n_evals; j0 += kTickEvalGroup) {
        const int nj = (int)min((int64_t)kTickEvalGroup, t.n_evals - j0);
        for (int jj = 0; jj < nj; ++jj) {
            const float *xw = s_win + (j0 + jj) * L + tid;   // window of evaluation j0 + jj: columns (j0 + jj) .. + T - 1, contiguous
            float acc[kPart];
#pragma unroll
            for (int h = 0; h < kPart; ++h) acc[h] = 0.0f;
            if (p.window_stat == FUSED_STAT_MINMAX) { acc[HP] = INFINITY; acc[HP + 1] = -INFINITY; }
#pragma unroll
            for (int k = 0; k < kW; ++k) {
                if (tid + k * kTickThreads < I) {
                    const float xv = xw[k * kTickThreads];
                    acc[0] = fmaf(xv, wa[k].x, acc[0]); acc[1] = fmaf(xv, wa[k].y, acc[1]); acc[2] = fmaf(xv, wa[k].z, acc[2]); acc[3] = fmaf(xv, wa[k].w, acc[3]);
                    if constexpr (HP == 8) {
                        acc[4] = fmaf(xv, wb[k].x, acc[4]); acc[5] = fmaf(xv, wb[k].y, acc[5]); acc[6] = fmaf(xv, wb[k].z, acc[6]); acc[7] = fmaf(xv, wb[k].w, acc[7]);
                    }
                    if (p.window_stat == FUSED_STAT_L2) acc[HP] = fmaf(xv, xv, acc[HP]);
                    else if (p.window_stat == FUSED_STAT_MINMAX) { acc[HP] = fminf(acc[HP], xv); acc[HP + 1] = fmaxf(acc[HP + 1], xv); }
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
                for (int h = 0; h < HP; ++h) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], d);
                if (p.window_stat == FUSED_STAT_MINMAX) {
                    acc[HP] = fminf(acc[HP], __shfl_xor_sync(0xffffffffu, acc[HP], d));
                    acc[HP + 1] = fmaxf(acc[HP + 1], __shfl_xor_sync(0xffffffffu, acc[HP + 1], d));
                } else acc[HP] += __shfl_xor_sync(0xffffffffu, acc[HP], d);
            }
            if (lane == 0) {
#pragma unroll
                for (int h = 0; h < kPart; ++h) s_part[(jj * 4 + warp) * kPart + h] = acc[h];
            }
        }
        __syncthreads();
        if (stamp) sub[4] = clock64();
        if (warp < nj) {
            const int64_t j = j0 + warp;
            const float *pq = s_part + warp * 4 * kPart;
            float acc[HP];
#pragma unroll
            for (int h = 0; h < HP; ++h) acc[h] = (pq[h] + pq[kPart + h]) + (pq[2 * kPart + h] + pq[3 * kPart + h]);
            float inv = 1.0f, beta = 0.0f;   // z = acc * inv + beta * V + B'
            if (p.window_stat == FUSED_STAT_L2) {                                         // x / sqrt(sum x^2)  (NeuralNet.swift:47-59); silence: NaN
                inv = 1.0f / sqrtf((pq[HP] + pq[kPart + HP]) + (pq[2 * kPart + HP] + pq[3 * kPart + HP]));
            } else if (p.window_stat == FUSED_STAT_MINMAX) {                              // NeuralNet.swift:69-96
                const float s0 = fminf(fminf(pq[HP], pq[kPart + HP]), fminf(pq[2 * kPart + HP], pq[3 * kPart + HP]));
                const float s1 = fmaxf(fmaxf(pq[HP + 1], pq[kPart + HP + 1]), fmaxf(pq[2 * kPart + HP + 1], pq[3 * kPart + HP + 1]));
                const float range = s1 - s0;
                if (0 == range) { inv = 0.0f; beta = -1.0f; }                             // flat window: every input becomes -1
                else { inv = 2.0f / range; beta = (0 - s0 - s1) / range; }
            }
            float a[kFusedMaxHidden], out[kFusedMaxOut];
#pragma unroll
            for (int h = 0; h < kFusedMaxHidden; ++h)
                a[h] = h < HP ? transfer_fast(p.tf[0], fmaf(acc[h < HP ? h : 0], inv, fmaf(beta, p.v[h], p.bprime[h]))) : 0.0f;
            network_tail(p, SYLDET_DETECT_ANY_OUTPUT, a, out);   // every lane: the same values
            if (stamp) sub[5] = clock64();
            if (lane == 0) {
                if (t.level_out) {   // output meter (Processor.swift:138, StatMax of Double(lastOutputs[0])); NaN never wins upstream either
                    const float v0 = out[0];
                    const int bits = __float_as_int(v0);
                    if (v0 == v0) atomicMax(t.level_out + ch, bits >= 0 ? bits : bits ^ 0x7fffffff);
                }
                if (t.packed) {
                    const uint4 r = make_uint4(__float_as_uint(out[0]), O > 1 ? __float_as_uint(out[1]) : 0u, O > 2 ? __float_as_uint(out[2]) : 0u, t.seq);
                    if constexpr (kResident)   // system scope: nothing ends here that would push a plain store out of the L2
                        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(t.packed + ch), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
                    else t.packed[ch] = r;
                } else
                    for (int o = 0; o < O; ++o) t.out[((int64_t)ch * t.n_evals + j) * O + o] = pick(out, o);
            }
        }
        if (j0 + kTickEvalGroup < t.n_evals) __syncthreads();   // s_part is reused
    }
    if (stamp) {
        ts[4] = clock64();
        for (int k = 0; k < 4; ++k) t.stamps[k] = ts[k];
        for (int k = 4; k < 10; ++k) t.stamps[k] = k == 4 ? ts[4] : ts[3];   // the generic tick's evaluation sub-phases do not exist here
        for (int k = 0; k < 6; ++k) t.stamps[10 + k] = sub[k];
    }
    if (t.flags && !t.packed) {   // publish: one flag per channel
        __threadfence_system();
        __syncthreads();
        if (tid == 0) *(volatile unsigned *)(t.flags + ch) = t.seq;
    }
}

template <int NFFT, int HP>
__global__ void __launch_bounds__(kTickThreads) stream_tick_fast_kernel(const FusedParams *__restrict__ pp, const StreamTick t, const TickGeom g,
                                                                        const float *__restrict__ window, const float2 *__restrict__ twiddle) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kHead = ((int)offsetof(FusedParams, w0) + 15) & ~15;
    TickInv<NFFT, HP> inv;
    const TickSmem sm = tick_smem_layout<NFFT, HP>(smem_raw, kHead, g);
    tick_run<NFFT, HP, false>(pp, t, g, inv, sm, window, twiddle);
}

// ---- resident tick: no launch on the latency path -------------------------------------------------------------------------
// The same tick, but the blocks stay on their SMs. The host posts the tick descriptor in pinned memory as 16-byte quads {3 payload
// words, sequence number} (a 16-byte store / PCIe read is indivisible, so every quad validates itself), the first quad last.
// Polling host memory is expensive in pollers, not in time (profiles/r02_poll_probe.txt: one polling thread answers in 2.5 us,
// every further poller of the line adds ~0.8 us), so ONE thread does it: block `n_channels` is the dispatcher - thread 0 polls the
// first quad, the block then fetches the message (one more PCIe round trip) and republishes it in a device-memory mailbox; the channel
// blocks poll that mailbox through the L2. The dispatcher alone decides to leave - when the host raises `quit`, or after
// `idle_cycles` without a message (a forgotten group releases its SMs; a cudaFree / synchronize elsewhere in the process stalls that
// long at most) - so a tick reaches every channel or none; the host sees `alive == 0` and starts the kernel again at the pending tick.
// The message holds only what changes from tick to tick (everything else is the StreamTick the kernel was started with):
//   word 0: n_cols | n_evals << 8 | n_marks << 16 | flags << 24 (bit 0: results go to `packed`)   word 1: n_staged
//   words 2-7: ring_pos, col0, eval0 (64-bit)   then marks[n_marks], and with the resampler rs_offset[n_marks], (rs_n_out | rs_out0 << 16)[n_marks]
// 12 words = 4 quads for the live shape (4 buffers per column, no resampler), 32 words at most.
constexpr int kPostWords = 8 + 3 * kStreamMaxMarks;
constexpr int kPostQuads = (kPostWords + 2) / 3;
static_assert(kPostQuads <= 32, "one quad of the tick message per lane of the dispatcher warp");
__host__ __device__ inline int tick_post_quads(int n_marks, bool rs) { return (8 + n_marks * (rs ? 3 : 1) + 2) / 3; }

__device__ __forceinline__ uint4 ld_volatile_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int NFFT, int HP>
__global__ void __launch_bounds__(kTickThreads) stream_tick_resident_kernel(const FusedParams *__restrict__ pp, const uint4 *post, const unsigned *quit,
                                                                            unsigned *alive, uint4 *mailbox, unsigned *leave, const StreamTick first,
                                                                            const TickGeom gmax, long long idle_cycles, int n_channels,
                                                                            const float *__restrict__ window, const float2 *__restrict__ twiddle) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kHead = ((int)offsetof(FusedParams, w0) + 15) & ~15;
    const int tid = threadIdx.x;
    __shared__ int s_go;
    unsigned expected = first.seq;
    if ((int)blockIdx.x == n_channels) {
        // ---- dispatcher: warp 0 polls the first four quads (one 64-byte line of host memory: the whole message of the live shape)
        if (tid >= 32) return;
        const int lane = tid;
        long long last = clock64();
        bool go = true;
        while (go) {
            uint4 v;
            int quads = 0;
            for (;;) {
                v = ld_volatile_v4(post + (lane < 4 ? lane : 0));
                const unsigned ok = __ballot_sync(0xffffffffu, v.w == expected);
                if (ok & 1u) {   // the first quad is written last: the message is complete in host memory
                    quads = tick_post_quads((__shfl_sync(0xffffffffu, v.x, 0) >> 16) & 0xff, first.rs_on != 0);
                    if ((ok & 0xfu) != 0xfu) v = ld_volatile_v4(post + (lane < quads ? lane : 0));   // this lane's read was the older one
                    break;
                }
                const bool stop = lane == 0 && (ld_volatile_u32(quit) != 0 || clock64() - last > idle_cycles);
                if (__any_sync(0xffffffffu, stop)) { go = false; break; }
            }
            if (!go) break;
            if (quads > 4 && lane >= 4 && lane < quads) v = ld_volatile_v4(post + lane);
            if (lane > 0 && lane < quads) __stcg(mailbox + lane, v);
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                asm volatile("st.release.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(mailbox), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                last = clock64();
            }
            ++expected;
        }
        if (lane == 0) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(leave), "r"(1u) : "memory");
            __threadfence_system();
            *(volatile unsigned *)alive = 0;
            __threadfence_system();
        }
        return;
    }
    // ---- one channel
    struct Mail {
        StreamTick t;
        TickGeom g;
        unsigned w[kPostQuads * 3];
    };
    static_assert(sizeof(Mail) <= 1024, "message area");
    Mail &m = *reinterpret_cast<Mail *>(smem_raw + kHead);
    const TickSmem sm = tick_smem_layout<NFFT, HP>(smem_raw, kHead + 1024, gmax);
    for (int i = tid; i < kHead / 16; i += kTickThreads) ptx::cp_async16(smem_raw + 16 * i, reinterpret_cast<const int4 *>(pp) + i);
    if (tid == 0) {
        m.t = first;
        m.g = gmax;
    }
    TickInv<NFFT, HP> inv;
    tick_load_inv<NFFT, HP>(inv, pp, gmax, window, twiddle, true, true);
    ptx::cp_async_wait_all();
    __syncthreads();
    for (;;) {
        if (tid == 0) {
            int go = 1;
            for (;;) {
                unsigned w, l;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(w) : "l"(&mailbox->w) : "memory");
                if (w == expected) break;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(l) : "l"(leave) : "memory");
                if (l != 0) { go = 0; break; }
            }
            s_go = go;
        }
        __syncthreads();
        if (!s_go) break;
        if (tid < kPostQuads) {
            const uint4 v = __ldcg(mailbox + tid);   // L2: written before the first quad was released (quads beyond the message: stale, unused)
            m.w[3 * tid] = v.x; m.w[3 * tid + 1] = v.y; m.w[3 * tid + 2] = v.z;
        }
        __syncthreads();
        if (tid == 0) {   // this tick's StreamTick: the constant part stays as the kernel was started with
            StreamTick &t = m.t;
            const unsigned *w = m.w;
            t.n_cols = w[0] & 0xff;
            t.n_evals = (w[0] >> 8) & 0xff;
            t.n_marks = (w[0] >> 16) & 0xff;
            t.packed = (w[0] >> 24) & 1 ? first.packed : nullptr;
            t.n_staged = (int)w[1];
            t.ring_pos = (int64_t)((unsigned long long)w[2] | (unsigned long long)w[3] << 32);
            t.col0 = (int64_t)((unsigned long long)w[4] | (unsigned long long)w[5] << 32);
            t.eval0 = (int64_t)((unsigned long long)w[6] | (unsigned long long)w[7] << 32);
            t.seq = expected;
            for (int k = 0; k < t.n_marks; ++k) {
                t.marks[k] = (int)w[8 + k];
                if (t.rs_on) {
                    t.rs_offset[k] = __uint_as_float(w[8 + t.n_marks + k]);
                    t.rs_n_out[k] = (int)(w[8 + 2 * t.n_marks + k] & 0xffff);
                    t.rs_out0[k] = (int)(w[8 + 2 * t.n_marks + k] >> 16);
                }
            }
            if (t.rs_on && expected != first.seq) {   // the two `last` arrays alternate from tick to tick
                const float *in = t.rs_last_out;
                t.rs_last_out = const_cast<float *>(t.rs_last_in);
                t.rs_last_in = in;
            }
            const int64_t span = t.n_cols > 0 ? (t.n_cols - 1) * (int64_t)gmax.hop + gmax.win_len : 0;
            const int64_t win = t.n_evals > 0 ? (t.col0 - t.eval0 + t.n_cols) * (int64_t)gmax.band : 0;
            m.g.span_floats = (int)((span + 3) & ~(int64_t)3);
            m.g.win_floats = (int)((win + 3) & ~(int64_t)3);
            m.g.stage_floats = (t.n_staged + 3) & ~3;
        }
        __syncthreads();
        tick_run<NFFT, HP, true>(pp, m.t, m.g, inv, sm, window, twiddle);
        __syncthreads();   // the message and the work areas are free again
        if (tid == 0) __threadfence_system();   // nothing ends here: without the fence the results sit in the L2 for hundreds of microseconds
        ++expected;
    }
}

constexpr size_t kTickFastSmemCap = 96 * 1024;

static TickGeom tick_geom(const FusedParams &p, const StreamTick &t) {
    TickGeom g{};
    g.win_len = p.win_len; g.hop = p.hop; g.gap = p.gap; g.k0 = p.k0; g.band = p.band; g.time_range = p.time_range;
    const int64_t span = t.n_cols > 0 ? (t.n_cols - 1) * (int64_t)p.hop + p.win_len : 0;
    const int64_t win = t.n_evals > 0 ? (t.col0 - t.eval0 + t.n_cols) * (int64_t)p.band : 0;
    g.span_floats = (int)((span + 3) & ~(int64_t)3);
    g.win_floats = (int)((win + 3) & ~(int64_t)3);
    g.stage_floats = (t.n_staged + 3) & ~3;
    return g;
}

static size_t tick_fast_smem(int fft_len, int hp, const TickGeom &g) {
    const size_t head = (offsetof(FusedParams, w0) + 15) & ~(size_t)15;
    const size_t part = (size_t)kTickEvalGroup * 4 * (hp + 2);   // even
    return head + ((size_t)g.span_floats + g.win_floats + g.stage_floats + part) * sizeof(float) +
           (size_t)4 * fused_group(fft_len) * fused_frame_pitch(fft_len) * sizeof(float2);
}

template <int NFFT, int HP>
cudaError_t launch_tick_one(const FusedParams *d_params, const StreamTick &t, const TickGeom &g, const float *window, const float2 *twiddle,
                            int n_channels, cudaStream_t stream) {
    static bool raised = false;   // once per instantiation: the attribute call is not free, and a tick is microseconds
    if (!raised) {
        cudaError_t e = cudaFuncSetAttribute(stream_tick_fast_kernel<NFFT, HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTickFastSmemCap);
        if (e != cudaSuccess) return e;
        raised = true;
    }
    stream_tick_fast_kernel<NFFT, HP><<<n_channels, kTickThreads, tick_fast_smem(NFFT, HP, g), stream>>>(d_params, t, g, window, twiddle);
    return cudaGetLastError();
}

template <int NFFT, int HP>
cudaError_t launch_resident_one(const FusedParams *d_params, const void *post, const unsigned *quit, unsigned *alive, void *mailbox, unsigned *leave,
                                const StreamTick &first, const TickGeom &gmax, long long idle_cycles, const float *window, const float2 *twiddle,
                                int n_channels, int sm_count, cudaStream_t stream) {
    auto kern = stream_tick_resident_kernel<NFFT, HP>;
    const size_t smem = tick_fast_smem(NFFT, HP, gmax) + 1024;   // + the message area
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTickFastSmemCap + 1024));
    if (e != cudaSuccess) return e;
    int per_sm = 0;   // the blocks wait for each other through memory: all of them (channels + dispatcher) have to be on an SM at once
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTickThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1 || n_channels + 1 > sm_count) return cudaErrorLaunchOutOfResources;
    kern<<<n_channels + 1, kTickThreads, smem, stream>>>(d_params, reinterpret_cast<const uint4 *>(post), quit, alive, reinterpret_cast<uint4 *>(mailbox), leave,
                                                         first, gmax, idle_cycles, n_channels, window, twiddle);
    return cudaGetLastError();
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

bool stream_tick_fast_supported(int fft_len, const FusedParams &p) {
    return fused_supports_fft(fft_len) && p.window_stat != FUSED_STAT_STD && p.band <= kFusedMaxBand;
}

bool stream_tick_fast_fits(int fft_len, int hp, const FusedParams &p, const StreamTick &t) {
    if (t.col0 < t.eval0) return false;
    return tick_fast_smem(fft_len, hp, tick_geom(p, t)) <= kTickFastSmemCap;
}

cudaError_t launch_stream_tick_fast(int fft_len, int hp, const FusedParams &host_params, const FusedParams *d_params, const StreamTick &t,
                                    const float *window, const float2 *twiddle, int n_channels, cudaStream_t stream) {
    const int cfg_fft = fft_len, cfg_hp = hp;
    const TickGeom g = tick_geom(host_params, t);
    SYLDET_FUSED_DISPATCH(launch_tick_one, d_params, t, g, window, twiddle, n_channels, stream)
}

size_t stream_tick_post_bytes() { return (size_t)kPostQuads * 16; }

void stream_tick_post_write(void *post, const StreamTick &t) {
    unsigned w[kPostQuads * 3] = {};
    w[0] = (unsigned)t.n_cols | (unsigned)t.n_evals << 8 | (unsigned)t.n_marks << 16 | (t.packed ? 1u : 0u) << 24;
    w[1] = (unsigned)t.n_staged;
    const unsigned long long q[3] = {(unsigned long long)t.ring_pos, (unsigned long long)t.col0, (unsigned long long)t.eval0};
    for (int k = 0; k < 3; ++k) { w[2 + 2 * k] = (unsigned)q[k]; w[3 + 2 * k] = (unsigned)(q[k] >> 32); }
    for (int k = 0; k < t.n_marks; ++k) {
        w[8 + k] = (unsigned)t.marks[k];
        if (t.rs_on) {
            std::memcpy(&w[8 + t.n_marks + k], &t.rs_offset[k], 4);
            w[8 + 2 * t.n_marks + k] = ((unsigned)t.rs_n_out[k] & 0xffff) | (unsigned)t.rs_out0[k] << 16;
        }
    }
    __m128i *dst = reinterpret_cast<__m128i *>(post);
    for (int q4 = tick_post_quads(t.n_marks, t.rs_on != 0) - 1; q4 >= 0; --q4)   // one indivisible 16-byte store per quad, the first quad last
        _mm_store_si128(dst + q4, _mm_set_epi32((int)t.seq, (int)w[3 * q4 + 2], (int)w[3 * q4 + 1], (int)w[3 * q4]));
    _mm_sfence();
}

bool stream_tick_resident_plan(int fft_len, int hp, const FusedParams &p, int stage_cap, TickGeom *gmax) {
    // the largest single-launch tick of the group: 4 columns, 4 evaluations, a full staging area
    TickGeom g{};
    g.win_len = p.win_len; g.hop = p.hop; g.gap = p.gap; g.k0 = p.k0; g.band = p.band; g.time_range = p.time_range;
    g.span_floats = (3 * p.hop + p.win_len + 3) & ~3;
    g.win_floats = ((p.time_range + 3) * p.band + 3) & ~3;
    g.stage_floats = (stage_cap + 3) & ~3;
    *gmax = g;
    return tick_fast_smem(fft_len, hp, g) <= kTickFastSmemCap;
}

bool stream_tick_resident_tick_fits(const TickGeom &gmax, const FusedParams &p, const StreamTick &t) {
    if (t.col0 < t.eval0 || t.n_cols > 255 || t.n_evals > 255 || t.n_marks > kStreamMaxMarks) return false;
    if (t.rs_on)
        for (int k = 0; k < t.n_marks; ++k)
            if (t.rs_n_out[k] > 0xffff || t.rs_out0[k] > 0xffff || t.rs_n_out[k] < 0) return false;
    const TickGeom g = tick_geom(p, t);
    return g.span_floats <= gmax.span_floats && g.win_floats <= gmax.win_floats && g.stage_floats <= gmax.stage_floats;
}

cudaError_t launch_stream_tick_resident(int fft_len, int hp, const FusedParams *d_params, const void *post, const unsigned *quit, unsigned *alive,
                                        void *mailbox, unsigned *leave, const StreamTick &first, const TickGeom &gmax, long long idle_cycles,
                                        const float *window, const float2 *twiddle, int n_channels, int sm_count, cudaStream_t stream) {
    const int cfg_fft = fft_len, cfg_hp = hp;
    SYLDET_FUSED_DISPATCH(launch_resident_one, d_params, post, quit, alive, mailbox, leave, first, gmax, idle_cycles, window, twiddle, n_channels,
                          sm_count, stream)
}

cudaError_t fused_max_blocks_per_sm(int fft_len, int hp, size_t smem, int *blocks) {
    const int cfg_fft = fft_len, cfg_hp = hp;
    SYLDET_FUSED_DISPATCH(occupancy_one, smem, blocks)
}

}  // namespace syldet
