// Tensor-core fused detection kernel (sm_100a): both contractions of the path run on tcgen05 as 3xTF32 products.
//
// Same path and citations as kernels_fused.cu (CSTFT.swift:280-337, SyllableDetector.swift:134-217, NeuralNet.swift:294-326,
// TrackDetector.swift:71-77). What changes is how the arithmetic is scheduled:
//
//  (1) band DFT.  The audio of a channel is viewed as a row-major matrix Y[row][hop] (row r = samples [r*hop, (r+1)*hop));
//      a frame starting at row c covers row c and the first W-hop samples of row c+1, so
//          X_c[k] = sum_n Y[c][n] B1[n][k] + sum_n Y[c+1][n] B2[n][k]      (B1/B2: halves of the windowed DFT matrix).
//      TMA (cp.async.bulk.tensor, SWIZZLE_128B/32B) lands 64-row tiles of Y in shared memory as the K-major B operand; the
//      DFT matrix [128 x 136] (rows: Re B1 | Im B1 | Re B2 | Im B2, 32 bins each) sits in TMEM as the A operand, split into
//      tf32 hi + lo; worker warps produce the audio lo part (x - tf32(x)); Ahi*Bhi + Ahi*Blo + Alo*Bhi accumulate
//      D[128 x 64] in TMEM (FP32, double buffered).
//  (2) layer 0.  Band magnitudes of 128 consecutive columns (hi/lo split, SWIZZLE_128B rows of 32 bins) are the A operand
//      of a second contraction against Wcat[(t,h)][f] = W'[t*L+f][h] (folded weights, hi/lo), giving per-column products
//      P[c][(t,h)] in TMEM; evaluation j then needs only the diagonal sum U_j[h] = sum_t P[j+t][(t,h)].
//  (3) epilogue (SIMT, one thread per evaluation): diagonal sum, window statistic from per-column partials, transfer
//      functions, remaining layers, reverse output maps, threshold test, event append.
// Roles: warps 0-7 workers (TMEM lane quadrant = warp % 4), warp 8 TMA producer, warp 9 MMA issuer + TMEM allocator.
#include <cuda.h>

#include "fused_epilogue.cuh"
#include "ptx_sm100.cuh"

namespace syldet {

namespace {

constexpr int kTcWorkers = 8;                  // worker warps
constexpr int kTcThreads = (kTcWorkers + 2) * 32;
constexpr int kTileRows = 64;                  // rows of Y per tile = N of the DFT MMA
constexpr int kTileFrames = kTileRows - 1;     // frames completed per tile (the last row only feeds the previous frame)
constexpr int kMainChunks = 4;                 // 32-float K chunks (SWIZZLE_128B)
constexpr int kTailCols = 8;                   // remaining K columns (SWIZZLE_32B)
constexpr int kKPad = kMainChunks * 32 + kTailCols;  // 136
constexpr int kMainBytes = kTileRows * 128;    // one main chunk
constexpr int kTailBytes = kTileRows * 32;
constexpr int kTileBytes = kMainChunks * kMainBytes + kTailBytes;  // 34 816 per hi (or lo) tile
constexpr int kGroup = 128;                    // columns per layer-0 MMA (its M)
constexpr int kMaxN0 = 56;                     // widest layer-0 product row (T * HP, padded to 16) with P double buffered
constexpr int kTmemCols = 512;
constexpr int kColAhi = 0, kColAlo = kKPad, kColD0 = 2 * kKPad, kColP0 = kColD0 + 2 * kTileRows;
static_assert(kColP0 + 2 * kMaxN0 <= kTmemCols, "TMEM budget");
constexpr int kStatRing = 256;

struct TcSmem {  // byte offsets from the 1024-byte aligned base
    static constexpr int hi0 = 0, hi1 = 35840, lo = 71680;          // audio tiles (34 816 rounded up to 1024)
    static constexpr int abuf = 107520;                             // [2 buffers][hi, lo][128 rows x 128 B]
    static constexpr int wcat = abuf + 4 * 16384;                   // [hi, lo][<= 56 rows x 128 B]
    static constexpr int xbuf = wcat + 2 * kMaxN0 * 128;            // [4 quadrants][64 frames][32 bins] float; aliased by pbuf
    static constexpr int carry = xbuf + 32768;                      // last T-1 product rows of the previous group
    static constexpr int colstat = carry + 4096;                    // float2[kStatRing] per-column statistic partials
    static constexpr int bars = colstat + kStatRing * 8;
    static constexpr int total = bars + 256;
    __host__ __device__ static constexpr int hi(int stage) { return stage ? hi1 : hi0; }
    __host__ __device__ static constexpr int a(int buf, int part) { return abuf + (buf * 2 + part) * 16384; }
};
static_assert(TcSmem::abuf % 1024 == 0 && TcSmem::wcat % 1024 == 0 && (kMaxN0 * 128) % 1024 == 0, "swizzle atoms need 1024-byte alignment");
static_assert(TcSmem::total + 1024 <= 227 * 1024, "shared memory budget");

// Linear walk over (unit, tile) pairs owned by this CTA; every role iterates the identical sequence.
struct TileWalk {
    int64_t unit, n_units;
    int tile, ntiles, ch, ncols, ne;
    int64_t e0;
    __device__ void load(const TcWork &w, int T) {
        if (unit >= n_units) return;
        ch = (int)(unit / w.chunks_per_channel);
        e0 = (unit - (int64_t)ch * w.chunks_per_channel) * w.chunk_evals;
        ne = (int)min(w.chunk_evals, w.evals_per_channel - e0);
        ncols = ne + T - 1;
        ntiles = (ncols + kTileFrames - 1) / kTileFrames;
        tile = 0;
    }
    __device__ void init(const TcWork &w, int T) {
        n_units = (int64_t)w.n_channels * w.chunks_per_channel;
        unit = blockIdx.x;
        load(w, T);
    }
    __device__ bool valid() const { return unit < n_units; }
    __device__ void next(const TcWork &w, int T) {
        if (++tile >= ntiles) {
            unit += gridDim.x;
            load(w, T);
        }
    }
    __device__ int first_row() const { return (int)e0 + tile * kTileFrames; }  // row index == column (frame) index
    __device__ int cols_before() const { return tile * kTileFrames; }
    __device__ int cols_after() const { return min(ncols, (tile + 1) * kTileFrames); }
    __device__ int groups_total() const { return (ncols + kGroup - 1) / kGroup; }
    // layer-0 groups (128 columns, the last one possibly partial) that become complete with this tile
    __device__ int groups_begin() const { return cols_before() / kGroup; }
    __device__ int groups_end() const { return tile == ntiles - 1 ? groups_total() : cols_after() / kGroup; }
};

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// byte offset of element (row, col) in a [rows][32 floats] SWIZZLE_128B tile (1024-byte aligned base)
__device__ __forceinline__ int sw128(int row, int col) { return row * 128 + ((((col >> 2) ^ row) & 7) << 4) + ((col & 3) << 2); }

struct Pending {  // layer-0 groups signalled to the MMA warp in the previous iteration, finalised in this one
    int count, g_begin, ch, ne, ncols;
    uint32_t gc_begin;
    int64_t e0;
};

template <int HP>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_detect_kernel(const __grid_constant__ FusedParams p, const TcWork w, const __grid_constant__ CUtensorMap tmap_main,
                 const __grid_constant__ CUtensorMap tmap_tail) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TcSmem::bars);
    uint64_t *full = bars, *ready = bars + 2, *stage_free = bars + 4, *tmem_full = bars + 6, *tmem_empty = bars + 8;
    uint64_t *a_ready = bars + 10, *p_full = bars + 12, *p_empty = bars + 14;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 16);
    float *xbuf = reinterpret_cast<float *>(smem + TcSmem::xbuf);
    float *pbuf = xbuf;  // alias: used between tiles only
    float *carry = reinterpret_cast<float *>(smem + TcSmem::carry);
    float2 *colstat = reinterpret_cast<float2 *>(smem + TcSmem::colstat);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.band, T = p.time_range;
    const int n0 = w.n0;              // layer-0 product row length (multiple of 16)
    const int ppitch = n0 + 1;        // pbuf / carry row pitch in floats

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&ready[i], kTcWorkers);
            ptx::mbar_init(&stage_free[i], 1);
            ptx::mbar_init(&tmem_full[i], 1);
            ptx::mbar_init(&tmem_empty[i], kTcWorkers);
            ptx::mbar_init(&a_ready[i], kTcWorkers);
            ptx::mbar_init(&p_full[i], 1);
            ptx::mbar_init(&p_empty[i], kTcWorkers);
        }
        ptx::fence_mbar_init();
    }
    if (warp == kTcWorkers + 1) {
        ptx::tmem_alloc(tmem_ptr, kTmemCols);
        ptx::tmem_relinquish();
    }
    // layer-0 weights -> shared memory, K-major SWIZZLE_128B rows of 32 bins (B operand of the second contraction)
    for (int i = tid; i < 2 * n0 * 32; i += kTcThreads) {
        const int part = i / (n0 * 32), r = (i / 32) % n0, c = i % 32;
        *reinterpret_cast<float *>(smem + TcSmem::wcat + part * kMaxN0 * 128 + sw128(r, c)) = __ldg((part ? w.wcat_lo : w.wcat_hi) + r * 32 + c);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // DFT matrix -> TMEM (A operand of the first contraction): lane m = matrix row, column = k, hi | lo
    if (warp < 4) {
        const int m = warp * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int part = 0; part < 2; ++part) {
            const float *src = (part ? w.dft_lo : w.dft_hi) + (size_t)m * kKPad;
            for (int kb = 0; kb < kKPad / 8; ++kb) {
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__ldg(src + kb * 8 + i));
                ptx::tmem_st_x8(lane_addr + (part ? kColAlo : kColAhi) + kb * 8, r);
            }
        }
        ptx::tc_wait_st();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    if (warp == kTcWorkers) {
        // ================================ TMA producer ================================================================
        if (lane == 0) {
            ptx::prefetch_tmap(&tmap_main);
            ptx::prefetch_tmap(&tmap_tail);
            TileWalk tw;
            tw.init(w, T);
            for (uint32_t it = 0; tw.valid(); ++it, tw.next(w, T)) {
                const int s = it & 1;
                ptx::mbar_wait(&stage_free[s], ((it >> 1) & 1) ^ 1);  // first use of each stage passes immediately
                unsigned char *dst = smem + TcSmem::hi(s);
                ptx::mbar_expect_tx(&full[s], kTileBytes);
                const int row = tw.first_row();
#pragma unroll
                for (int j = 0; j < kMainChunks; ++j) ptx::tma_load_3d(dst + j * kMainBytes, &tmap_main, j * 32, row, tw.ch, &full[s]);
                ptx::tma_load_3d(dst + kMainChunks * kMainBytes, &tmap_tail, kMainChunks * 32, row, tw.ch, &full[s]);
            }
        }
    } else if (warp == kTcWorkers + 1) {
        // ================================ MMA issuer ==================================================================
        if (lane == 0) {
            constexpr uint32_t idesc_dft = ptx::idesc_tf32(128, kTileRows);
            const uint32_t idesc_l0 = ptx::idesc_tf32(kGroup, n0);
            const uint32_t lo = ptx::smem_addr(smem + TcSmem::lo);
            const uint32_t wc_hi = ptx::smem_addr(smem + TcSmem::wcat), wc_lo = wc_hi + kMaxN0 * 128;
            uint32_t gc = 0;  // layer-0 groups issued so far (global over units)
            auto issue_l0 = [&]() {
                const int ab = gc & 1;
                const uint32_t ph = (gc >> 1) & 1;
                ptx::mbar_wait(&a_ready[ab], ph);       // magnitudes of the group written and fenced
                ptx::mbar_wait(&p_empty[ab], ph ^ 1);   // product buffer drained by the epilogue
                ptx::tc_fence_after();
                const uint32_t d = tmem_base + kColP0 + ab * kMaxN0;
                const uint32_t a_hi = ptx::smem_addr(smem + TcSmem::a(ab, 0)), a_lo = ptx::smem_addr(smem + TcSmem::a(ab, 1));
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a = pass == 1 ? a_lo : a_hi, b = pass == 2 ? wc_lo : wc_hi;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        ptx::mma_tf32_ss(d, ptx::smem_desc_kmajor(a + ks * 32, 1024, 2), ptx::smem_desc_kmajor(b + ks * 32, 1024, 2), idesc_l0, acc);
                        acc = 1;
                    }
                }
                ptx::mma_commit(&p_full[ab]);
                ++gc;
            };
            TileWalk tw;
            tw.init(w, T);
            int n_pending = 0;
            for (uint32_t it = 0; tw.valid(); ++it, tw.next(w, T)) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                ptx::mbar_wait(&ready[s], ph);            // hi landed (TMA) and lo written (workers)
                ptx::mbar_wait(&tmem_empty[s], ph ^ 1);   // accumulator s drained by the epilogue
                ptx::tc_fence_after();
                const uint32_t d = tmem_base + kColD0 + s * kTileRows;
                const uint32_t hi = ptx::smem_addr(smem + TcSmem::hi(s));
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a = tmem_base + (pass == 2 ? kColAlo : kColAhi);
                    const uint32_t b = pass == 1 ? lo : hi;
#pragma unroll
                    for (int j = 0; j < kMainChunks; ++j)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            ptx::mma_tf32_ts(d, a + j * 32 + ks * 8, ptx::smem_desc_kmajor(b + j * kMainBytes + ks * 32, 1024, 2), idesc_dft, acc);
                            acc = 1;
                        }
                    ptx::mma_tf32_ts(d, a + kMainChunks * 32, ptx::smem_desc_kmajor(b + kMainChunks * kMainBytes, 256, 6), idesc_dft, acc);
                }
                ptx::mma_commit(&tmem_full[s]);
                ptx::mma_commit(&stage_free[s]);
                // layer-0 products of the groups the PREVIOUS tile completed (their magnitudes are written while this DFT runs)
                for (int k = 0; k < n_pending; ++k) issue_l0();
                n_pending = tw.groups_end() - tw.groups_begin();
            }
            for (int k = 0; k < n_pending; ++k) issue_l0();
        }
    } else {
        // ================================ workers =====================================================================
        const int quad = warp & 3, half = warp >> 2;
        auto worker_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kTcWorkers * 32) : "memory"); };
        auto split_lo = [&](uint32_t it) {  // lo = x - tf32_trunc(x) for the whole tile (layout-agnostic: same offsets in both buffers)
            const int s = it & 1;
            ptx::mbar_wait(&full[s], (it >> 1) & 1);
            const float4 *hi4 = reinterpret_cast<const float4 *>(smem + TcSmem::hi(s));
            float4 *lo4 = reinterpret_cast<float4 *>(smem + TcSmem::lo);
#pragma unroll 3
            for (int i = tid; i < kTileBytes / 16; i += kTcWorkers * 32) {
                const float4 v = hi4[i];
                float4 o;
                o.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                o.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                o.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                o.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                lo4[i] = o;
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&ready[s]);
        };

        // ---- finalise the evaluations covered by layer-0 group g of a unit ---------------------------------------------
        auto finalize = [&](const Pending &pd, int k) {
            const int g = pd.g_begin + k;
            const uint32_t gcg = pd.gc_begin + k;
            const int ab = gcg & 1;
            ptx::mbar_wait(&p_full[ab], (gcg >> 1) & 1);
            ptx::tc_fence_after();
            // P(g) [128 lanes x n0] -> pbuf rows [T-1, T-1+128)
            {
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + kColP0 + ab * kMaxN0;
                float *dst = pbuf + (T - 1 + quad * 32 + lane) * ppitch;
                for (int cc = half * (n0 / 2); cc < (half + 1) * (n0 / 2); cc += 8) {
                    uint32_t r[8];
                    ptx::tmem_ld_x8(taddr + cc, r);
                    ptx::tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) dst[cc + i] = __uint_as_float(r[i]);
                }
                ptx::tc_fence_before();
            }
            // per-column statistic partials from the magnitudes (hi + lo is exact): sum of squares or (min, max)
            if (p.window_stat != FUSED_STAT_NONE && tid < kGroup) {
                const int col = g * kGroup + tid;
                if (col < pd.ncols) {
                    const unsigned char *ahi = smem + TcSmem::a(ab, 0), *alo = smem + TcSmem::a(ab, 1);
                    float s0 = p.window_stat == FUSED_STAT_L2 ? 0.0f : INFINITY, s1 = -INFINITY;
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const int off = tid * 128 + (((c4 ^ tid) & 7) << 4);
                        const float4 h4 = *reinterpret_cast<const float4 *>(ahi + off), l4 = *reinterpret_cast<const float4 *>(alo + off);
                        const float m[4] = {h4.x + l4.x, h4.y + l4.y, h4.z + l4.z, h4.w + l4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (c4 * 4 + i < L) {
                                if (p.window_stat == FUSED_STAT_L2) s0 = fmaf(m[i], m[i], s0);
                                else { s0 = fminf(s0, m[i]); s1 = fmaxf(s1, m[i]); }
                            }
                    }
                    colstat[col & (kStatRing - 1)] = make_float2(s0, s1);
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_empty[ab]);
            if (g > 0)
                for (int i = tid; i < (T - 1) * n0; i += kTcWorkers * 32) pbuf[(i / n0) * ppitch + i % n0] = carry[(i / n0) * ppitch + i % n0];
            worker_sync();
            // evaluations whose T columns end inside this group
            const int jb = max(0, g * kGroup - (T - 1));
            const int je = min(pd.ne, (g + 1) * kGroup - (T - 1));
            float *out_base = w.all_out ? w.all_out + ((int64_t)pd.ch * w.out_evals_per_channel + w.eval_offset + pd.e0) * p.n_out : nullptr;
            for (int qb = warp * 32; qb < je - jb; qb += kTcWorkers * 32) {
                const int j = jb + qb + lane;
                float out[kFusedMaxOut];
                bool hit = false;
                if (j < je) {
                    const float *prow = pbuf + (j - (g * kGroup - (T - 1))) * ppitch;
                    float acc[HP];
#pragma unroll
                    for (int h = 0; h < HP; ++h) acc[h] = 0.0f;
                    float s0 = p.window_stat == FUSED_STAT_L2 ? 0.0f : INFINITY, s1 = -INFINITY;
                    for (int t = 0; t < T; ++t) {
#pragma unroll
                        for (int h = 0; h < HP; ++h) acc[h] += prow[t * ppitch + t * HP + h];
                        if (p.window_stat != FUSED_STAT_NONE) {
                            const float2 cs = colstat[(j + t) & (kStatRing - 1)];
                            if (p.window_stat == FUSED_STAT_L2) s0 += cs.x;
                            else { s0 = fminf(s0, cs.x); s1 = fmaxf(s1, cs.y); }
                        }
                    }
                    float alpha_div, beta;
                    bool constant_input;
                    stat_to_affine(p.window_stat, s0, s1, alpha_div, beta, constant_input);
                    hit = finish_eval<HP>(p, w.detect_rule, acc, alpha_div, beta, constant_input, out);
                    if (out_base) {
                        float *o = out_base + (int64_t)j * p.n_out;
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
                            w.sink.events[idx] = DevEvent{pd.ch, 0, w.eval_offset + pd.e0 + j};
#pragma unroll
                            for (int i = 0; i < kFusedMaxOut; ++i)
                                if (i < p.n_out) w.sink.outputs[idx * p.n_out + i] = out[i];
                        }
                    }
                }
            }
            worker_sync();
            // keep the last T-1 product rows for the next group of this unit
            for (int i = tid; i < (T - 1) * n0; i += kTcWorkers * 32)
                carry[(i / n0) * ppitch + i % n0] = pbuf[(kGroup + i / n0) * ppitch + i % n0];
            worker_sync();
        };

        TileWalk cur;
        cur.init(w, T);
        if (cur.valid()) split_lo(0);
        Pending pend{};
        uint32_t gc = 0;        // layer-0 groups signalled so far (global over units)
        uint32_t gc_unit = 0;   // value of gc when the current unit started
        for (uint32_t it = 0; cur.valid(); ++it) {
            TileWalk nxt = cur;
            nxt.next(w, T);
            const int s = it & 1;
            if (cur.tile == 0) gc_unit = gc;
            ptx::mbar_wait(&tmem_full[s], (it >> 1) & 1);  // DFT of this tile done; the lo buffer is free again
            ptx::tc_fence_after();
            if (nxt.valid()) split_lo(it + 1);              // the next tile's MMAs overlap everything below

            for (int k = 0; k < pend.count; ++k) finalize(pend, k);
            pend.count = 0;

            // ---- D (TMEM) -> xbuf[quadrant][frame][bin] ----------------------------------------------------------
            const int frames = cur.cols_after() - cur.cols_before();
            {
                uint32_t r[32];
                ptx::tmem_ld_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + kColD0 + s * kTileRows + half * 32, r);
                ptx::tc_wait_ld();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty[s]);
                float *dst = xbuf + (quad * kTileRows + half * 32) * 32 + lane;
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[j * 32] = __uint_as_float(r[j]);
            }
            worker_sync();
            // ---- X_c = (Re1[c] + Re2[c+1]) + i (Im1[c] + Im2[c+1]); |X| of the band -> layer-0 A operand (hi, lo) ----------
            for (int c = warp; c < frames; c += kTcWorkers) {
                const int col = cur.cols_before() + c;
                const int ab = (gc_unit + (col >> 7)) & 1, row = col & (kGroup - 1);
                const float re = xbuf[(0 * kTileRows + c) * 32 + lane] + xbuf[(2 * kTileRows + c + 1) * 32 + lane];
                const float im = xbuf[(1 * kTileRows + c) * 32 + lane] + xbuf[(3 * kTileRows + c + 1) * 32 + lane];
                float mag = sqrt_fast(re * re + im * im);
                if (p.scaling != SYLDET_SCALING_LINEAR) mag = scale_value(mag, p.scaling);
                if (lane >= L) mag = 0.0f;
                const float hi = tf32_rna(mag);
                const int off = sw128(row, lane);
                *reinterpret_cast<float *>(smem + TcSmem::a(ab, 0) + off) = hi;
                *reinterpret_cast<float *>(smem + TcSmem::a(ab, 1) + off) = mag - hi;
                if (w.debug_band && lane < L) w.debug_band[((int64_t)cur.ch * w.debug_cols + cur.e0 + col) * L + lane] = mag;
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            // ---- hand completed groups to the MMA warp; they are finalised in the next iteration -------------------------
            const int gb = cur.groups_begin(), ge = cur.groups_end();
            if (ge > gb) {
                if (lane == 0)
                    for (int g = gb; g < ge; ++g) ptx::mbar_arrive(&a_ready[(gc_unit + g) & 1]);
                pend.count = ge - gb;
                pend.g_begin = gb;
                pend.gc_begin = gc_unit + gb;
                pend.ch = cur.ch;
                pend.ne = cur.ne;
                pend.ncols = cur.ncols;
                pend.e0 = cur.e0;
                gc = gc_unit + ge;
            }
            worker_sync();  // xbuf is reused by the next finalize / tile
            cur = nxt;
        }
        for (int k = 0; k < pend.count; ++k) finalize(pend, k);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kTcWorkers + 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

size_t tc_smem_bytes(const FusedParams &) { return 1024 + TcSmem::total; }
int tc_tile_frames() { return kTileFrames; }
int tc_k_pad() { return kKPad; }
int tc_group_cols() { return kGroup; }
int tc_max_n0() { return kMaxN0; }
bool tc_layout_fits(int time_range, int n0) {
    return n0 <= kMaxN0 && (time_range - 1) * (n0 + 1) * 4 <= 4096 && (time_range - 1 + kGroup) * (n0 + 1) * 4 <= 32768 &&
           kGroup + time_range <= kStatRing;
}

cudaError_t launch_tc(int hp, int grid, size_t smem, const FusedParams &p, const TcWork &w, const void *tmap_main, const void *tmap_tail,
                      cudaStream_t stream) {
    const CUtensorMap &tm = *static_cast<const CUtensorMap *>(tmap_main);
    const CUtensorMap &tt = *static_cast<const CUtensorMap *>(tmap_tail);
    cudaError_t e;
    if (hp == 4) {
        e = cudaFuncSetAttribute(tc_detect_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tc_detect_kernel<4><<<grid, kTcThreads, smem, stream>>>(p, w, tm, tt);
    } else {
        e = cudaFuncSetAttribute(tc_detect_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tc_detect_kernel<8><<<grid, kTcThreads, smem, stream>>>(p, w, tm, tt);
    }
    return cudaGetLastError();
}

}  // namespace syldet
