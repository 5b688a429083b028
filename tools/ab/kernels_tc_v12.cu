// Tensor-core fused detection kernel (sm_100a): both contractions of the path run on tcgen05 as 3xTF32 products, the
// SIMT work is split over warp-specialised roles that only meet through mbarriers.
//
// Same path and citations as kernels_fused.cu (CSTFT.swift:280-337, SyllableDetector.swift:134-217, NeuralNet.swift:294-326,
// TrackDetector.swift:71-77). What changes is how the arithmetic is scheduled:
//
//  (1) band DFT.  The audio of a channel is viewed as a row-major matrix Y[row][hop] (row r = samples [r*hop, (r+1)*hop));
//      a frame starting at row c covers row c and the first W-hop samples of row c+1, so
//          X_c[k] = sum_n Y[c][n] B1[n][k] + sum_n Y[c+1][n] B2[n][k]      (B1/B2: halves of the windowed DFT matrix).
//      TMA (cp.async.bulk.tensor, SWIZZLE_128B/32B) lands 64-row tiles of Y in shared memory as the K-major B operand; the
//      DFT matrix [128 x 136] (rows: Re B1 | Im B1 | Re B2 | Im B2, 32 bins each) sits in TMEM as the A operand, split into
//      tf32 hi + lo; splitter warps produce the audio lo part (x - tf32(x)); Ahi*Bhi + Alo*Bhi + Ahi*Blo accumulate
//      D[128 x 64] in TMEM (FP32, double buffered).  A tile completes 63 frames (its last row only closes frame 62).
//  (2) layer 0.  The band magnitudes of the tile's 63 columns (hi/lo split, SWIZZLE_128B rows of 32 bins) are the A operand
//      (M = 64) of a second contraction against Wcat[(t,h)][f] = W'[t*L+f][h] (folded weights, hi/lo), giving per-column
//      products P[c][(t,h)] in TMEM; evaluation j then needs only the diagonal sum U_j[h] = sum_t P[j+t][(t,h)].
//  (3) epilogue (SIMT): diagonal sum over a shared-memory ring of product rows, window statistic from per-column
//      partials, transfer functions, remaining layers, reverse output maps, threshold test, event append.
//
// Roles (22 warps with two evaluator groups, one CTA per SM, persistent over (channel, chunk) units; every role walks the same tile sequence):
//   warp 0        TMA producer                      full[s] <- hi_free[s]
//   warp 1        MMA issuer + TMEM allocator       DFT(it): full, (tmem_empty: implied), lo_ready -> hi_free, tmem_full, lo_free
//                                                   layer0(it-1): a_ready, p_empty -> p_full, a_free
//   then          evaluators (F): TC_GROUPS (2) groups of four warps (one per TMEM lane quadrant) that take tiles in turn:
//                                                   p_full, ring_ready[other] -> product ring -> p_empty, ring_ready[own];
//                                                   diagonal sum (T x LDS.128), window statistic, network tail, events
//   then 8        spectrum warps (D)                tmem_full -> D -> registers -> tmem_empty; two shuffle rounds -> |X| ->
//                                                   layer-0 A operand + per-column statistic partials -> a_ready
//   then 4        splitters (S)                     full, lo_free -> lo tile -> lo_ready, hi_free
// Every role is a serial chain per tile and runs at ~0.1 IPC per warp (profiles/): the roles overlap through double buffers, and
// the longest chain (the evaluators': ~530 dependent warp instructions per tile) is split over the groups so that no role
// needs more than one tile period per tile.
#include <cuda.h>
#include <cuda_fp16.h>

#include "fused_epilogue.cuh"
#include "ptx_sm100.cuh"

namespace syldet {

namespace {

constexpr int kFGroup = 4;                     // warps per evaluator group (one per TMEM lane quadrant)
// kF16: the two correction products of the 3xTF32 band DFT (Alo*Bhi + Ahi*Blo) as ONE K-concatenated kind::f16 pass
//   [fp16(Alo) | fp16(Ahi * 2^-11)] * [fp16(x) ; fp16((x - tf32(x)) * 2^11)]       (17 MMAs instead of 34 per tile)
// The fp16 tile has the byte geometry of the fp32 one (4 x [64 rows x 128 B] + [64 rows x 32 B]); K order: x(n < 128) |
// lo(n < 128) | x(n >= 128) | lo(n >= 128). Error terms are scaled by 2^-11, so fp16's 11 bits keep the sum at fp32 level
// for |x| in [6e-5, 65504]; quieter samples carry an absolute error of 2^-25 * 2^-11 each (documented in DESIGN.md).
#ifndef TC_GROUPS
#define TC_GROUPS 2
#endif
constexpr int kNumGroups = TC_GROUPS;          // evaluator groups; group g takes the tiles with it % kNumGroups == g
constexpr int kWarpTma = 0, kWarpMma = 1, kWarpF0 = 2, kNumF = kFGroup * kNumGroups, kWarpD0 = kWarpF0 + kNumF, kNumD = 8,
              kWarpS0 = kWarpD0 + kNumD, kNumS = 4;
constexpr int kTcThreads = (kWarpS0 + kNumS) * 32;  // 832
static_assert(kWarpF0 % 4 == 2 && kWarpD0 % 4 == 2, "quadrant / half assignment below assumes these starts");
constexpr int kTileRows = 64;                  // rows of Y per tile = N of the DFT MMA
constexpr int kTileFrames = kTileRows - 1;     // frames completed per tile
constexpr int kMainChunks = 4;                 // 32-float K chunks (SWIZZLE_128B)
constexpr int kTailCols = 8;                   // remaining K columns (SWIZZLE_32B)
constexpr int kKPad = kMainChunks * 32 + kTailCols;  // 136
constexpr int kMainBytes = kTileRows * 128;    // one main chunk
constexpr int kTailBytes = kTileRows * 32;
constexpr int kTileBytes = kMainChunks * kMainBytes + kTailBytes;  // 34 816 per hi (or lo) tile
constexpr int kMaxN0 = 56;                     // widest layer-0 product row (T * HP, padded to 16) with P double buffered
constexpr int kTmemCols = 512;
constexpr int kColAhi = 0, kColAlo = kKPad, kColD0 = 2 * kKPad, kColP0 = kColD0 + 2 * kTileRows;
static_assert(kColP0 + 2 * kMaxN0 <= kTmemCols, "TMEM budget");
constexpr int kPRing = kNumGroups * kTileFrames + 22;   // product-row ring (rows = columns): one tile per group in flight + the T-1 (<= 21) rows before them
constexpr int kStatRing = 512;                 // per-column statistic ring: [2 planes][kStatRing][4 bin quarters] float
constexpr int kBarF = 2;                       // named barriers of the F groups: kBarF and kBarF + 1
constexpr int kEvCap = 96;                     // shared-memory event buffer per F group (flushed with one global atomic)
static_assert(kBarF + kNumGroups <= 16 && kNumGroups <= 4, "named barriers / barrier block layout");

struct TcSmem {  // byte offsets from the 1024-byte aligned base
    static constexpr int hi0 = 0, hi1 = kTileBytes, lo = 2 * kTileBytes;   // audio tiles (34 816 B = 34 swizzle atoms each)
    static constexpr int abuf = 3 * kTileBytes;                     // [2 buffers][hi, lo][64 rows x 128 B]
    static constexpr int wcat = abuf + 4 * 8192;                    // [hi, lo][<= 56 rows x 128 B]
    static constexpr int pbuf = wcat + 2 * kMaxN0 * 128;            // [kPRing][ppitch] float; everything after it is placed at run time
    // then: float4 colstat[planes][kStatRing] | event meta int4[groups][kEvCap] | event outputs float[groups][kEvCap][n_out] |
    //       barriers (256 B) | optionally the second lo tile (1024-byte aligned) when it fits
    // planes: 1 (sum of squares), 2 for the min/max statistic
    __host__ __device__ static constexpr int ppitch(int np) { return ((((np + 7) >> 3) << 1) | 1) << 2; }  // whole 8-float chunks + 1: an odd number of float4
    __host__ __device__ static constexpr int colstat(int np) { return pbuf + kPRing * ppitch(np) * 4; }
    __host__ __device__ static constexpr int evmeta(int np, int planes) { return colstat(np) + planes * kStatRing * 16; }
    __host__ __device__ static constexpr int evout(int np, int planes) { return evmeta(np, planes) + kNumGroups * kEvCap * 16; }
    __host__ __device__ static constexpr int bars(int np, int n_out, int planes) { return evout(np, planes) + kNumGroups * kEvCap * n_out * 4; }
    __host__ __device__ static constexpr int lo1(int np, int n_out, int planes) { return (bars(np, n_out, planes) + 256 + 1023) & ~1023; }
    __host__ __device__ static constexpr int total(int np, int n_out, int planes, int lo_stages) {
        return lo_stages == 2 ? lo1(np, n_out, planes) + kTileBytes : bars(np, n_out, planes) + 256;
    }
    __host__ __device__ static constexpr int hi(int stage) { return stage ? hi1 : hi0; }
    __host__ __device__ static constexpr int a(int buf, int part) { return abuf + (buf * 2 + part) * 8192; }
};
static_assert(kTileBytes % 1024 == 0, "tiles are whole swizzle atoms");
static_assert(TcSmem::abuf % 1024 == 0 && TcSmem::wcat % 1024 == 0 && (kMaxN0 * 128) % 1024 == 0, "swizzle atoms need 1024-byte alignment");
static_assert(TcSmem::pbuf % 16 == 0, "alignment");

// Linear walk over (unit, tile) pairs owned by this CTA; every role iterates the identical sequence.
struct TileWalk {
    int units_left;       // units this CTA still has to start (including the current one)
    int ch, chunk;        // current unit = (channel, chunk of evaluations); advanced by gridDim.x units without divisions
    int tile, ntiles, ncols, ne;
    int64_t e0;
    __device__ __forceinline__ void load(const TcWork &w, int T) {
        e0 = (int64_t)chunk * w.chunk_evals;
        ne = (int)min(w.chunk_evals, w.evals_per_channel - e0);
        ncols = ne + T - 1;
        ntiles = (ncols + kTileFrames - 1) / kTileFrames;
        tile = 0;
    }
    __device__ __forceinline__ void advance(const TcWork &w, int by) {
        chunk += by;
        while (chunk >= w.chunks_per_channel) {
            chunk -= w.chunks_per_channel;
            ++ch;
        }
    }
    __device__ __forceinline__ void init(const TcWork &w, int T) {
        const int n_units = w.n_channels * w.chunks_per_channel;
        units_left = n_units > (int)blockIdx.x ? (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
        ch = 0;
        chunk = 0;
        advance(w, (int)blockIdx.x);
        if (units_left > 0) load(w, T);
    }
    __device__ __forceinline__ bool valid() const { return units_left > 0; }
    __device__ __forceinline__ void next(const TcWork &w, int T) {
        if (++tile >= ntiles) {
            if (--units_left > 0) {
                advance(w, (int)gridDim.x);
                load(w, T);
            }
        }
    }
    __device__ __forceinline__ int first_row() const { return (int)e0 + tile * kTileFrames; }  // row index == column (frame) index
    __device__ __forceinline__ int cols_before() const { return tile * kTileFrames; }
    __device__ __forceinline__ int frames() const { return min(ncols - tile * kTileFrames, kTileFrames); }
};

__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// byte offset of element (row, col) in a [rows][32 floats] SWIZZLE_128B tile (1024-byte aligned base)
__device__ __forceinline__ int sw128(int row, int col) { return row * 128 + ((((col >> 2) ^ row) & 7) << 4) + ((col & 3) << 2); }
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Optional per-role cycle accounting (kOn: the SYLDET_TC_TIMING instantiation): slot += cycles spent in a wait / in the whole
// role loop. The production instantiation compiles to the bare waits.
template <bool kOn>
struct RoleTimer {
    long long *dst;
    long long acc[6];
    long long t_start;
    bool relaxed;   // this role has slack: its waits may back off (ptx::mbar_wait_relaxed)
    __device__ RoleTimer(long long *base, int first_slot, bool relaxed_waits = true) : dst(nullptr), acc{0, 0, 0, 0, 0, 0}, t_start(0), relaxed(relaxed_waits) {
        if constexpr (kOn) {
            dst = base + blockIdx.x * 32 + first_slot;
            t_start = clock64();
        }
    }
    __device__ __forceinline__ long long now() const {
        if constexpr (kOn) return clock64();
        return 0;
    }
    __device__ __forceinline__ void add(int k, long long t0) {
        if constexpr (kOn) acc[k] += clock64() - t0;
    }
    __device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity, int k) {
        const long long t0 = now();
        if (relaxed) ptx::mbar_wait_relaxed(bar, parity);
        else ptx::mbar_wait(bar, parity);
        add(k, t0);
    }
    __device__ __forceinline__ void sync(int id, int n, int k) {
        const long long t0 = now();
        bar_sync(id, n);
        add(k, t0);
    }
    __device__ void flush(bool writer) {
        if constexpr (kOn) {
            if (writer) {
                acc[5] = clock64() - t_start;
                for (int k = 0; k < 6; ++k) dst[k] = acc[k];
            }
        }
    }
};

// An evaluator group empties its shared-memory event buffer: all threads of the group; one global atomic for the whole batch.
// Out of line (one copy, away from the hot loop).
__device__ __noinline__ void flush_group_events(EventSink sink, const int4 *ev_meta, const float *ev_out, int *ev_count,
                                                unsigned long long *ev_base, int n_ev, int ft, int bar_f, int n_out) {
    if (ft == 0) *ev_base = atomicAdd(sink.count, (unsigned long long)n_ev);
    bar_sync(bar_f, kFGroup * 32);
    const unsigned long long base = *ev_base;
    for (int e = ft; e < n_ev; e += kFGroup * 32) {
        const unsigned long long idx = base + e;
        if (idx < sink.capacity) {
            const int4 m = ev_meta[e];
            sink.events[idx] = DevEvent{m.x, 0, (int64_t)(((unsigned long long)(unsigned)m.w << 32) | (unsigned)m.z)};
            for (int k = 0; k < n_out; ++k) sink.outputs[idx * n_out + k] = ev_out[e * n_out + k];
        }
    }
    bar_sync(bar_f, kFGroup * 32);
    if (ft == 0) *ev_count = 0;
}

// Rare shapes (more than two layers, several outputs): kept out of line so that the evaluators' hot loop stays compact.
__device__ __noinline__ bool network_tail_cold(const FusedParams &p, int detect_rule, float (&a)[kFusedMaxHidden], float (&out)[kFusedMaxOut]) {
    return network_tail(p, detect_rule, a, out);
}

// kFast: the shape of the reference's sample network is known at compile time (l2normalize window statistic, tansig hidden
// layer, one purelin output, one reverse output map), which strips the run-time dispatch from the evaluators' dependent chain.
template <int HP, bool kScaled, bool kTiming, bool kFast, bool kF16>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_detect_kernel(const __grid_constant__ FusedParams p, const TcWork w, const __grid_constant__ CUtensorMap tmap_main,
                 const __grid_constant__ CUtensorMap tmap_tail) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    // 1024-byte aligned base (swizzle atoms); plain pointer arithmetic so that the compiler keeps the shared address space
    unsigned char *smem = smem_dyn + ((1024u - (ptx::smem_addr(smem_dyn) & 1023u)) & 1023u);
    const int L = p.band, T = p.time_range;
    const int window_stat = kFast ? (int)FUSED_STAT_L2 : p.window_stat;
    const int tf0 = kFast ? (int)SYLDET_TF_TANSIG : p.tf[0], tf1 = kFast ? (int)SYLDET_TF_PURELIN : p.tf[1];
    const int n_op = kFast ? 1 : p.n_op, n_out = kFast ? 1 : p.n_out;
    const bool one_output_tail = kFast || (p.n_layers == 2 && p.n_out == 1);
    const int n0 = w.n0;                           // layer-0 product row length (multiple of 16)
    const int np = T * HP;                         // its meaningful prefix
    const int ppitch = TcSmem::ppitch(np);         // product ring pitch in floats
    const int planes = window_stat == FUSED_STAT_MINMAX ? 2 : 1;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TcSmem::bars(np, p.n_out, planes));
    uint64_t *full = bars, *hi_free = bars + 2, *lo_ready = bars + 4, *lo_free = bars + 29, *tmem_full = bars + 6, *tmem_empty = bars + 8;   // lo_*: [2]
    // lo tile(s): with two, stage = it & 1 like the hi tiles and the splitters never wait for pass 3 of the previous tile
    const int lo_stages = w.lo_stages;
    const int lo1_off = TcSmem::lo1(np, p.n_out, planes);
    uint64_t *a_ready = bars + 10, *a_free = bars + 12, *p_full = bars + 14, *p_empty = bars + 16;
    uint64_t *ring_ready = bars + 18;                                         // [groups <= 4]: an F group has written its tile's product rows
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 22);
    int *ev_counts = reinterpret_cast<int *>(bars + 23);                      // [groups <= 4] events waiting in shared memory, per F group
    unsigned long long *ev_bases = reinterpret_cast<unsigned long long *>(bars + 25);  // [groups <= 4]
    float *pbuf = reinterpret_cast<float *>(smem + TcSmem::pbuf);
    float *colstat = reinterpret_cast<float *>(smem + TcSmem::colstat(np));   // sum of squares | minimum, then the maximum plane

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&hi_free[i], 1 + kNumS);
            ptx::mbar_init(&tmem_full[i], 1);
            ptx::mbar_init(&tmem_empty[i], kNumD);
            ptx::mbar_init(&a_ready[i], kNumD);
            ptx::mbar_init(&a_free[i], 1);
            ptx::mbar_init(&p_full[i], 1);
            ptx::mbar_init(&p_empty[i], kFGroup);
        }
        for (int i = 0; i < kNumGroups; ++i) {
            ptx::mbar_init(&ring_ready[i], kFGroup);
            ev_counts[i] = 0;
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&lo_ready[i], kNumS);
            ptx::mbar_init(&lo_free[i], 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == kWarpMma) {
        ptx::tmem_alloc(tmem_ptr, kTmemCols);
        ptx::tmem_relinquish();
    }
    // layer-0 weights -> shared memory, K-major SWIZZLE_128B rows of 32 bins (B operand of the second contraction)
    for (int i = tid; i < 2 * n0 * 32; i += kTcThreads) {
        const int part = i / (n0 * 32), r = (i / 32) % n0, c = i % 32;
        *reinterpret_cast<float *>(smem + TcSmem::wcat + part * kMaxN0 * 128 + sw128(r, c)) = __ldg((part ? w.wcat_lo : w.wcat_hi) + r * 32 + c);
    }
    // rows of the layer-0 A operand that no tile writes (row 63, short last tiles) must hold finite values
    for (int i = tid; i < 4 * 8192 / 16; i += kTcThreads) reinterpret_cast<float4 *>(smem + TcSmem::abuf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // DFT matrix -> TMEM (A operand of the first contraction): column = k, hi | lo; lane 32*quad + lane takes the row of
    // (part = lane >> 3, bin = 8*quad + (lane & 7)), which puts the four parts of a bin into one warp of the spectrum role
    if (warp >= kWarpF0 && warp < kWarpF0 + 4) {
        const int quad = warp & 3;
        const int m = (lane >> 3) * 32 + quad * 8 + (lane & 7);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int part = 0; part < 2; ++part) {
            const float *src = (part ? (kF16 ? reinterpret_cast<const float *>(w.dft16) : w.dft_lo) : w.dft_hi) + (size_t)m * kKPad;
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

    if (warp == kWarpTma) {
        // ================================ TMA producer ================================================================
        if (ptx::elect_one()) {
            ptx::prefetch_tmap(&tmap_main);
            ptx::prefetch_tmap(&tmap_tail);
            TileWalk tw;
            tw.init(w, T);
            RoleTimer<kTiming> tm(w.debug_timing, 0);
            for (uint32_t it = 0; tw.valid(); ++it, tw.next(w, T)) {
                const int s = it & 1;
                tm.wait(&hi_free[s], ((it >> 1) & 1) ^ 1, 0);  // first use of each stage passes immediately
                unsigned char *dst = smem + TcSmem::hi(s);
                ptx::mbar_expect_tx(&full[s], kTileBytes);
                const int row = tw.first_row();
#pragma unroll
                for (int j = 0; j < kMainChunks; ++j) ptx::tma_load_3d(dst + j * kMainBytes, &tmap_main, j * 32, row, tw.ch, &full[s]);
                ptx::tma_load_3d(dst + kMainChunks * kMainBytes, &tmap_tail, kMainChunks * 32, row, tw.ch, &full[s]);
            }
            tm.flush(true);
        }
    } else if (warp == kWarpMma) {
        // ================================ MMA issuer ==================================================================
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_dft = ptx::idesc_tf32(128, kTileRows);
            const uint32_t idesc_l0 = ptx::idesc_tf32(64, n0);
            const uint32_t lo_a = ptx::smem_addr(smem + TcSmem::lo), lo_b = ptx::smem_addr(smem + lo1_off);
            const uint32_t wc_hi = ptx::smem_addr(smem + TcSmem::wcat), wc_lo = wc_hi + kMaxN0 * 128;
            // one K sweep of the band DFT: 4 SWIZZLE_128B chunks of 4 k-steps + the 8-column SWIZZLE_32B tail (rolled: the
            // instruction stream of every role has to stay small, the roles share the instruction cache)
            auto dft_pass = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t acc) {
                uint64_t desc = ptx::smem_desc_kmajor(b, 1024, 2);
#pragma unroll 1
                for (int j = 0; j < kMainChunks; ++j, a += 32, desc += kMainBytes >> 4) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        ptx::mma_tf32_ts(d, a + ks * 8, desc + ks * 2, idesc_dft, acc);
                        acc = 1;
                    }
                }
                ptx::mma_tf32_ts(d, a, ptx::smem_desc_kmajor(b + kMainChunks * kMainBytes, 256, 6), idesc_dft, 1);
            };
            // the fp16 correction pass: same descriptors and column steps (32 B of K per instruction), 16 k each
            constexpr uint32_t idesc_f16 = ptx::idesc_f16(128, kTileRows);
            auto corr_pass = [&](uint32_t d, uint32_t a, uint32_t b) {
                uint64_t desc = ptx::smem_desc_kmajor(b, 1024, 2);
#pragma unroll 1
                for (int j = 0; j < kMainChunks; ++j, a += 32, desc += kMainBytes >> 4) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) ptx::mma_f16_ts(d, a + ks * 8, desc + ks * 2, idesc_f16, 1);
                }
                ptx::mma_f16_ts(d, a, ptx::smem_desc_kmajor(b + kMainChunks * kMainBytes, 256, 6), idesc_f16, 1);
            };
            RoleTimer<kTiming> tm(w.debug_timing, 6, false);   // the MMA issuer is the pacemaker: it polls
            auto issue_l0 = [&](uint32_t jt) {  // per-column layer-0 products of tile jt
                const int ab = jt & 1;
                const uint32_t ph = (jt >> 1) & 1;
                tm.wait(&a_ready[ab], ph, 3);       // magnitudes written and fenced
                tm.wait(&p_empty[ab], ph ^ 1, 4);   // product buffer drained by the evaluators
                ptx::tc_fence_after();
                const uint32_t d = tmem_base + kColP0 + ab * kMaxN0;
                const uint32_t a_hi = ptx::smem_addr(smem + TcSmem::a(ab, 0)), a_lo = ptx::smem_addr(smem + TcSmem::a(ab, 1));
                uint32_t acc = 0;
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass) {
                    const uint64_t da = ptx::smem_desc_kmajor(pass == 1 ? a_lo : a_hi, 1024, 2), db = ptx::smem_desc_kmajor(pass == 2 ? wc_lo : wc_hi, 1024, 2);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        ptx::mma_tf32_ss(d, da + ks * 2, db + ks * 2, idesc_l0, acc);
                        acc = 1;
                    }
                }
                ptx::mma_commit(&p_full[ab]);
                ptx::mma_commit(&a_free[ab]);
#ifdef TC_EXP_SERIAL_P   // experiment: no MMA queued while the evaluators read P
                ptx::mbar_wait(&p_empty[ab], ph);
#endif
            };
            TileWalk tw;
            tw.init(w, T);
            uint32_t it = 0;
            for (; tw.valid(); ++it, tw.next(w, T)) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                tm.wait(&full[s], ph, 0);             // hi landed (TMA)
                // accumulator s is free: the spectrum warps read it for tile it-2 before they arrived on a_ready(it-2), which
                // this thread waited for when it issued layer 0 of that tile (previous iteration) - no separate wait needed
#ifdef TC_WAIT_TMEM_EMPTY
                tm.wait(&tmem_empty[s], ph ^ 1, 1);
#endif
                ptx::tc_fence_after();
                const uint32_t d = tmem_base + kColD0 + s * kTileRows;
                const uint32_t hi = ptx::smem_addr(smem + TcSmem::hi(s));
                dft_pass(d, tmem_base + kColAhi, hi, 0);
                if constexpr (!kF16) dft_pass(d, tmem_base + kColAlo, hi, 1);
                ptx::mma_commit(&hi_free[s]);             // the MMA side is done with hi[s]
                const int ls = lo_stages == 2 ? s : 0;
                const uint32_t lo_use = lo_stages == 2 ? (it >> 1) : it;   // uses of this lo buffer so far
                tm.wait(&lo_ready[ls], lo_use & 1, 2);   // lo written (splitters)
                ptx::tc_fence_after();
                if constexpr (kF16) corr_pass(d, tmem_base + kColAlo, ls ? lo_b : lo_a);
                else dft_pass(d, tmem_base + kColAhi, ls ? lo_b : lo_a, 1);
                ptx::mma_commit(&tmem_full[s]);
                ptx::mma_commit(&lo_free[ls]);
                if (it > 0) issue_l0(it - 1);             // its magnitudes were written while this tile's DFT was queued
            }
            if (it > 0) issue_l0(it - 1);
            tm.flush(true);
        }
    } else if (warp < kWarpD0) {
        // ================================ evaluators (F) ==============================================================
        // Groups of four warps (one warp per TMEM lane quadrant); group g takes the tiles with it % groups == g, reading the
        // layer-0 accumulator P[it & 1]. The accumulator is an M = 64 tile: product row c (= column c of the tile) sits in lane
        // 32*(c/16) + c%16, so lanes 0-15 of a warp own one row each: they move it to the product ring and then evaluate the
        // network whose newest column is c. The ring is shared by the groups: ring_ready[g] says "group g wrote its tile's rows".
        const int quad = warp & 3, grp = (warp - kWarpF0) >> 2;
        const int ft = (warp - kWarpF0 - grp * kFGroup) * 32 + lane;   // thread index inside the group
        const int bar_f = kBarF + grp;
        int *ev_count = ev_counts + grp;
        unsigned long long *ev_base = ev_bases + grp;
        int4 *ev_meta = reinterpret_cast<int4 *>(smem + TcSmem::evmeta(np, planes)) + grp * kEvCap;   // (channel, -, eval lo, eval hi)
        float *ev_out = reinterpret_cast<float *>(smem + TcSmem::evout(np, planes)) + grp * kEvCap * n_out;
        const int c = quad * 16 + (lane & 15);              // column of the tile this thread owns
        const bool owner = lane < 16;
        const int nchunks = (np + 7) >> 3;                  // 8-column chunks of a product row
        TileWalk tw;
        tw.init(w, T);
        int gcol = 0;                                       // product-ring position of the tile's first column (mod kPRing)
        uint32_t scol = 0;                                  // columns seen so far, all units (statistic ring position)
        RoleTimer<kTiming> tm(w.debug_timing, 12);
        auto flush_events = [&](int n_ev) { flush_group_events(w.sink, ev_meta, ev_out, ev_count, ev_base, n_ev, ft, bar_f, n_out); };
        int turn = 0;                                       // it % kNumGroups
        for (uint32_t it = 0; tw.valid(); ++it, tw.next(w, T), turn = turn + 1 == kNumGroups ? 0 : turn + 1) {
            const int ab = it & 1;                          // layer-0 accumulator of this tile
            const int frames = tw.frames();
            if (turn != grp) {                              // another group's tile: only keep the ring positions in step
                gcol += frames;
                if (gcol >= kPRing) gcol -= kPRing;
                scol += frames;
                continue;
            }
            // the group of tile it-1 has written that tile (history rows), which also proves it finished the evaluations of tile
            // it-1-groups, the newest tile whose ring rows this tile may reuse (this group's own tile it-groups is done as well)
            if (it > 0) tm.wait(&ring_ready[grp == 0 ? kNumGroups - 1 : grp - 1], ((it - 1) / kNumGroups) & 1, 0);
            // only now: a group sees every other use of an accumulator, and a parity wait is only meaningful for the phase that
            // is pending; tile it-1 copied => tile it-2 (the previous use of P[ab]) copied => the pending phase is this tile's
            tm.wait(&p_full[ab], (it >> 1) & 1, 0);
            ptx::tc_fence_after();
            const long long t_f0 = tm.now();
            {   // P row -> product ring, four 8-column chunks at a time: loads in flight, one wait, then the stores
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + kColP0 + ab * kMaxN0;
                int row = gcol + c;
                if (row >= kPRing) row -= kPRing;
                float4 *dst = reinterpret_cast<float4 *>(pbuf + row * ppitch);
                const bool store = owner && c < frames;
#pragma unroll
                for (int q0 = 0; q0 < kMaxN0 / 8; q0 += 4) {
                    if (q0 < nchunks) {
                        uint32_t r[4][8];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q0 + q < nchunks) ptx::tmem_ld_x8(taddr + (q0 + q) * 8, r[q]);
                        ptx::tc_wait_ld();
                        if (store) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (q0 + q < nchunks) {  // the ring pitch covers whole chunks
                                    dst[2 * (q0 + q)] = make_float4(__uint_as_float(r[q][0]), __uint_as_float(r[q][1]), __uint_as_float(r[q][2]), __uint_as_float(r[q][3]));
                                    dst[2 * (q0 + q) + 1] = make_float4(__uint_as_float(r[q][4]), __uint_as_float(r[q][5]), __uint_as_float(r[q][6]), __uint_as_float(r[q][7]));
                                }
                        }
                    }
                }
                tm.add(3, t_f0);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&p_empty[ab]);
                    ptx::mbar_arrive(&ring_ready[grp]);
                }
            }
            tm.sync(bar_f, kFGroup * 32, 1);
            const long long t_f1 = tm.now();
            // evaluation whose newest column is column c of this tile: unit-local index j
            const int j = tw.cols_before() - (T - 1) + c;
#ifdef TC_EXP_SKIP_EVAL
            const bool valid = false;
#else
            const bool valid = c < frames && j >= 0;
#endif
            // Lanes l and l + 16 share evaluation c: each sums half of the T columns, one shuffle round combines them, both run
            // the (cheap) network tail and lane l stores.
            bool hit = false;
            float out[kFusedMaxOut];
            {
                float acc[HP];
#pragma unroll
                for (int h = 0; h < HP; ++h) acc[h] = 0.0f;
                float s0 = window_stat == FUSED_STAT_L2 ? 0.0f : INFINITY, s1 = -INFINITY;
                if (valid) {
                    const int th = (T + 1) >> 1, t0 = owner ? 0 : th, t1 = owner ? th : T;
                    int row = gcol + c - (T - 1) + t0;                     // ring position of this lane's first column
                    if (row < 0) row += kPRing;
                    else if (row >= kPRing) row -= kPRing;
                    uint32_t col = scol + (uint32_t)(c - (T - 1) + t0);    // same column in the statistic ring
                    const float *pt = pbuf + t0 * HP + row * ppitch;       // advances by one ring row and one (t, :) block per step
#pragma unroll 2
                    for (int t = t0; t < t1; ++t, ++col) {
                        const float4 *prow = reinterpret_cast<const float4 *>(pt);
                        pt += ppitch + HP;
                        if (++row == kPRing) { row = 0; pt -= kPRing * ppitch; }
                        const float4 v0 = prow[0];
                        acc[0] += v0.x; acc[1] += v0.y; acc[2] += v0.z; acc[3] += v0.w;
                        if constexpr (HP == 8) {
                            const float4 v1 = prow[1];
                            acc[4] += v1.x; acc[5] += v1.y; acc[6] += v1.z; acc[7] += v1.w;
                        }
                        const float4 *cs = reinterpret_cast<const float4 *>(colstat) + (col & (kStatRing - 1));
                        const float4 ca = cs[0];            // the four bin quarters of the column
                        if (window_stat == FUSED_STAT_L2) s0 += (ca.x + ca.y) + (ca.z + ca.w);
                        else if (window_stat == FUSED_STAT_MINMAX) {
                            const float4 cb = cs[kStatRing];
                            s0 = fminf(s0, fminf(fminf(ca.x, ca.y), fminf(ca.z, ca.w)));
                            s1 = fmaxf(s1, fmaxf(fmaxf(cb.x, cb.y), fmaxf(cb.z, cb.w)));
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < HP; ++h) acc[h] += __shfl_xor_sync(0xffffffffu, acc[h], 16);
                {
                    const float o0 = __shfl_xor_sync(0xffffffffu, s0, 16), o1 = __shfl_xor_sync(0xffffffffu, s1, 16);
                    if (window_stat == FUSED_STAT_L2) s0 += o0;
                    else { s0 = fminf(s0, o0); s1 = fmaxf(s1, o1); }
                }
                float inv = 1.0f, beta = 0.0f;  // z = acc * inv + beta * V + B'
                if (window_stat == FUSED_STAT_L2) {            // x / sqrt(sum x^2)  (NeuralNet.swift:47-59); silence: 0 * inf = NaN
                    inv = rcp_fast(sqrt_fast(s0));
                } else if (window_stat == FUSED_STAT_MINMAX) {  // x * 2/range + (-mn-mx)/range  (NeuralNet.swift:69-96)
                    const float range = s1 - s0;
                    if (0 == range) { inv = 0.0f; beta = -1.0f; }  // flat window: every input becomes -1
                    else { inv = 2.0f / range; beta = (0 - s0 - s1) / range; }
                }
                float a[kFusedMaxHidden];
#pragma unroll
                for (int h = 0; h < kFusedMaxHidden; ++h)
                    a[h] = h < HP ? transfer_fast(tf0, fmaf(acc[h < HP ? h : 0], inv, fmaf(beta, p.v[h], p.bprime[h]))) : 0.0f;
                tm.add(4, t_f1);
                float *o = w.all_out + ((int64_t)tw.ch * w.out_evals_per_channel + w.eval_offset + tw.e0 + j) * n_out;
                const bool store = valid && owner && w.all_out != nullptr;
                if (one_output_tail) {   // the common shape: hidden layer -> one output (NeuralNet.swift:310-323)
                    float sacc = p.rest_b[0];
#pragma unroll
                    for (int h = 0; h < HP; ++h) sacc = fmaf(p.rest_w[h], a[h], sacc);
                    float v = transfer_fast(tf1, sacc);
#pragma unroll 1
                    for (int k = 0; k < n_op; ++k)  // reverse maps in index order
                        v = (v + (0 - p.op_y[k])) / p.op_gain[k * kFusedMaxOut] + p.op_xoff[k * kFusedMaxOut];
                    hit = v >= p.thr_f[0];            // == (double)v >= thr (TrackDetector.swift:72); NaN -> false
                    out[0] = v;
                    if (store) o[0] = v;
                } else {
                    float a_mem[kFusedMaxHidden], out_mem[kFusedMaxOut];   // copies: only this branch touches local memory
#pragma unroll
                    for (int h = 0; h < kFusedMaxHidden; ++h) a_mem[h] = a[h];
                    hit = network_tail_cold(p, w.detect_rule, a_mem, out_mem);
#pragma unroll
                    for (int k = 0; k < kFusedMaxOut; ++k) out[k] = out_mem[k];
                    if (store) {
#pragma unroll 1
                        for (int k = 0; k < n_out; ++k) o[k] = pick(out, k);
                    }
                }
                hit = hit && valid && owner;
            }
            const unsigned hits = __ballot_sync(0xffffffffu, hit);
            if (hits) {   // append to the shared-memory event buffer (room for a whole tile is guaranteed by the flush rule)
                int base = 0;
                if (lane == 0) base = atomicAdd(ev_count, __popc(hits));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) {
                    const int e = base + __popc(hits & ((1u << lane) - 1));
                    const int64_t ev = w.eval_offset + tw.e0 + j;
                    ev_meta[e] = make_int4(tw.ch, 0, (int)(unsigned)(ev & 0xffffffffll), (int)(ev >> 32));
#pragma unroll 1
                    for (int k = 0; k < n_out; ++k) ev_out[e * n_out + k] = pick(out, k);
                }
            }
            tm.sync(bar_f, kFGroup * 32, 2);  // this group's next tile overwrites ring rows that this tile's evaluations read
            const int n_ev = *ev_count;
            if (n_ev > kEvCap - kTileFrames) flush_events(n_ev);
            gcol += frames;
            if (gcol >= kPRing) gcol -= kPRing;
            scol += frames;
        }
        {
            const int n_ev = *ev_count;
            if (n_ev > 0) flush_events(n_ev);
        }
        tm.flush(ft == 0 && grp == 0);
    } else if (warp < kWarpS0) {
        // ================================ spectrum warps (D) ===========================================================
        // TMEM lane 32*quad + lane holds DFT-matrix row (part = lane >> 3: Re B1 | Im B1 | Re B2 | Im B2, bin = 8*quad + (lane & 7)),
        // so the four parts of a bin sit in one warp and combine through two shuffle rounds; nothing goes through shared memory:
        //   (1) X[c] = P1[c] + P2[c+1]: B1 lanes (bit 4 clear) take columns 4g, 4g+1, B2 lanes take 4g+2, 4g+3 (xor 16)
        //   (2) re^2 + im^2: Re lanes (bit 3 clear) keep the columns of even g, Im lanes those of odd g           (xor 8)
        // after which every lane owns 8 magnitudes of its bin: columns col0 + 8*i + u. A store instruction then covers rows
        // r, r+2, r+4, r+6 x 8 bins = all 32 banks of the SWIZZLE_128B operand.
        const int quad = warp & 3, dw = warp - kWarpD0, half = dw >> 2;
        const bool up = (lane & 16) != 0, im = (lane & 8) != 0;
        const int bin = quad * 8 + (lane & 7);
        const bool in_band = bin < L;
        const int col0 = half * 32 + (up ? 2 : 0) + (im ? 4 : 0);
        int a_off[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) a_off[u] = (col0 + u) * 128 + ((((bin >> 2) ^ (col0 + u)) & 7) << 4) + ((bin & 3) << 2);
        const int stat_col = col0 + 8 * ((lane & 7) >> 1) + (lane & 1);   // the column whose statistic ends up in this lane
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + kColD0 + half * 32;
        TileWalk tw;
        tw.init(w, T);
        uint32_t gcol = 0;
        RoleTimer<kTiming> tm(w.debug_timing, 18);
        for (uint32_t it = 0; tw.valid(); ++it, tw.next(w, T)) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int frames = tw.frames();
            tm.wait(&tmem_full[s], ph, 0);
            ptx::tc_fence_after();
            // One load of the warp's 32 (+1) columns, then straight-line code: the shuffle chains of the eight column groups are
            // independent, and a warp needs that many in flight to hide the latency of the (busy) shared-memory/shuffle pipe.
            uint32_t r[33];
            const long long t_d0 = tm.now();
            {
                uint32_t r32[32];
                ptx::tmem_ld_x32(taddr0 + s * kTileRows, r32);
                r[32] = 0;                                      // column 64 does not exist: frame 63 is never complete in this tile
                if (half == 0) ptx::tmem_ld_x1(taddr0 + s * kTileRows + 32, r[32]);
                ptx::tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = r32[j];
            }
            tm.add(4, t_d0);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[s]);
            // ---- (1) X_c = (Re1[c] + Re2[c+1]) + i (Im1[c] + Im2[c+1]) ------------------------------------------------------
            float x[8][2];
#pragma unroll
            for (int g = 0; g < 8; ++g)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float send = __uint_as_float(up ? r[4 * g + u + 1] : r[4 * g + 2 + u]);
                    const float own = __uint_as_float(up ? r[4 * g + 3 + u] : r[4 * g + u]);
                    x[g][u] = own + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            // ---- (2) |X| of the band (plain multiplies and add, as the reference computes it) -------------------------------
            float mag[8];                                       // e = 2*i + u: column col0 + 8*i + u
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float sq_a = __fmul_rn(x[2 * i][u], x[2 * i][u]), sq_b = __fmul_rn(x[2 * i + 1][u], x[2 * i + 1][u]);
                    const float other = __shfl_xor_sync(0xffffffffu, im ? sq_a : sq_b, 8);
                    float v;
                    if constexpr (kScaled) v = scale_value_nl(sqrt_fast(__fadd_rn(im ? sq_b : sq_a, other)), p.scaling);
                    else v = sqrt_fast_ftz(__fadd_rn(im ? sq_b : sq_a, other));
                    mag[2 * i + u] = in_band ? v : 0.0f;
                }
            // ---- window statistic: per-column partial over this warp's 8 bins; the evaluators combine the four quadrants -----
            if (window_stat != FUSED_STAT_NONE) {
                // eight values x eight lanes -> lane b holds the total of value b: exchange half of the values per round
                auto reduce8 = [&](float (&q)[8], auto op) {
                    const bool b4 = (lane & 4) != 0, b2 = (lane & 2) != 0, b1 = (lane & 1) != 0;
                    float h4[4], h2[2];
#pragma unroll
                    for (int k = 0; k < 4; ++k) h4[k] = op(b4 ? q[k + 4] : q[k], __shfl_xor_sync(0xffffffffu, b4 ? q[k] : q[k + 4], 4));
#pragma unroll
                    for (int k = 0; k < 2; ++k) h2[k] = op(b2 ? h4[k + 2] : h4[k], __shfl_xor_sync(0xffffffffu, b2 ? h4[k] : h4[k + 2], 2));
                    return op(b1 ? h2[1] : h2[0], __shfl_xor_sync(0xffffffffu, b1 ? h2[0] : h2[1], 1));
                };
                float q[8];
                float *cs = colstat + (((gcol + (uint32_t)stat_col) & (kStatRing - 1)) << 2) + quad;
                if (window_stat == FUSED_STAT_L2) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) q[e] = mag[e] * mag[e];
                    const float tot = reduce8(q, [](float a, float b) { return a + b; });
                    if (stat_col < frames) cs[0] = tot;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) q[e] = in_band ? mag[e] : INFINITY;
                    const float mn = reduce8(q, [](float a, float b) { return fminf(a, b); });
#pragma unroll
                    for (int e = 0; e < 8; ++e) q[e] = in_band ? mag[e] : -INFINITY;
                    const float mx = reduce8(q, [](float a, float b) { return fmaxf(a, b); });
                    if (stat_col < frames) {
                        cs[0] = mn;
                        cs[kStatRing * 4] = mx;
                    }
                }
            }
            // ---- magnitudes -> layer-0 A operand (hi, lo) -----------------------------------------------------------------------
            tm.wait(&a_free[s], ph ^ 1, 2);                 // layer 0 of tile it-2 has read this A buffer
            {
                unsigned char *a_hi = smem + TcSmem::a(s, 0), *a_lo = smem + TcSmem::a(s, 1);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int col = col0 + 8 * i + u;
                        if (col < frames) {
                            const float v = mag[2 * i + u], vh = tf32_trunc(v);
                            *reinterpret_cast<float *>(a_hi + a_off[u] + i * 1024) = vh;
                            *reinterpret_cast<float *>(a_lo + a_off[u] + i * 1024) = v - vh;
                            if (w.debug_band && in_band)
                                w.debug_band[((int64_t)tw.ch * w.debug_cols + tw.e0 + tw.cols_before() + col) * L + bin] = v;
                        }
                    }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&a_ready[s]);
            gcol += frames;
        }
        tm.flush(dw == 0 && lane == 0);
    } else {
        // ================================ splitters (S) ================================================================
        const int st = (warp - kWarpS0) * 32 + lane;
        constexpr int kPerThread = kTileBytes / 16 / (kNumS * 32);
        static_assert(kPerThread * kNumS * 32 * 16 == kTileBytes, "tile size must divide over the splitters");
        TileWalk tw;
        tw.init(w, T);
        RoleTimer<kTiming> tm(w.debug_timing, 24);
        for (uint32_t it = 0; tw.valid(); ++it, tw.next(w, T)) {
            const int s = it & 1;
            tm.wait(&full[s], (it >> 1) & 1, 0);
            const int ls = lo_stages == 2 ? s : 0;
            const uint32_t lo_use = lo_stages == 2 ? (it >> 1) : it;
            tm.wait(&lo_free[ls], (lo_use & 1) ^ 1, 1);  // pass 3 of the previous user of this lo buffer has read it
            if constexpr (kF16) {   // fp32 tile -> fp16 tile [x | (x - tf32(x)) * 2^11]. A warp instruction takes rows r0, r0+4, r0+8, r0+12 (one 128 B row
                // per quarter warp): the loads are whole rows and the 8-byte stores of the four rows fall on disjoint banks.
                const unsigned char *hi_b = smem + TcSmem::hi(s);
                unsigned char *b16 = smem + (ls ? lo1_off : TcSmem::lo);
                const int sw = st >> 5, phys = lane & 7, rq = sw + 4 * (lane >> 3);   // row = rq + 16 * (kk & 3), chunk = kk >> 2
                auto convert = [](const float4 &v, uint2 &hx, uint2 &hl) {
                    const __half2 x01 = __floats2half2_rn(v.x, v.y), x23 = __floats2half2_rn(v.z, v.w);
                    const __half2 l01 = __floats2half2_rn((v.x - tf32_trunc(v.x)) * 2048.0f, (v.y - tf32_trunc(v.y)) * 2048.0f);
                    const __half2 l23 = __floats2half2_rn((v.z - tf32_trunc(v.z)) * 2048.0f, (v.w - tf32_trunc(v.w)) * 2048.0f);
                    hx = make_uint2(*reinterpret_cast<const uint32_t *>(&x01), *reinterpret_cast<const uint32_t *>(&x23));
                    hl = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
                };
#pragma unroll
                for (int b0 = 0; b0 < 16; b0 += 8) {
                    float4 v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int kk = b0 + k, r = rq + 16 * (kk & 3), c = kk >> 2;
                        v[k] = *reinterpret_cast<const float4 *>(hi_b + c * kMainBytes + r * 128 + phys * 16);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int kk = b0 + k, r = rq + 16 * (kk & 3), c = kk >> 2;
                        const int u = phys ^ (r & 7);                      // logical 16-byte unit: samples 32*c + 4*u .. + 3
                        const int unit16 = 4 * (c & 1) + (u >> 1);         // their place in the fp16 row (64 k per 128 B)
                        unsigned char *dst = b16 + (c >> 1) * kMainBytes + r * 128 + ((unit16 ^ (r & 7)) << 4) + (u & 1) * 8;
                        uint2 hx, hl;
                        convert(v[k], hx, hl);
                        *reinterpret_cast<uint2 *>(dst) = hx;
                        *reinterpret_cast<uint2 *>(dst + 2 * kMainBytes) = hl;
                    }
                }
                {   // tail: samples 128 .. 135 of row r (SWIZZLE_32B: 16-byte unit ^= bit 2 of the row)
                    const int r = st >> 1, ph = st & 1, flip = (r >> 2) & 1, u = ph ^ flip;
                    const float4 v = *reinterpret_cast<const float4 *>(hi_b + kMainChunks * kMainBytes + r * 32 + ph * 16);
                    unsigned char *dst = b16 + kMainChunks * kMainBytes + r * 32 + u * 8;
                    uint2 hx, hl;
                    convert(v, hx, hl);
                    *reinterpret_cast<uint2 *>(dst + ((0 ^ flip) << 4)) = hx;   // x: k 256 .. 263 = logical unit 0
                    *reinterpret_cast<uint2 *>(dst + ((1 ^ flip) << 4)) = hl;   // lo: k 264 .. 271 = logical unit 1
                }
            } else {
            const float4 *hi4 = reinterpret_cast<const float4 *>(smem + TcSmem::hi(s)) + st;
            float4 *lo4 = reinterpret_cast<float4 *>(smem + (ls ? lo1_off : TcSmem::lo)) + st;
            // lo = x - tf32_trunc(x); layout-agnostic: same offsets in both buffers. Two batches, loads in flight before stores.
            constexpr int kBatch = (kPerThread + 1) / 2;
#pragma unroll
            for (int b0 = 0; b0 < kPerThread; b0 += kBatch) {
                float4 v[kBatch];
#pragma unroll
                for (int k = 0; k < kBatch; ++k)
                    if (b0 + k < kPerThread) v[k] = hi4[(b0 + k) * kNumS * 32];
#pragma unroll
                for (int k = 0; k < kBatch; ++k)
                    if (b0 + k < kPerThread)
                        lo4[(b0 + k) * kNumS * 32] = make_float4(v[k].x - tf32_trunc(v[k].x), v[k].y - tf32_trunc(v[k].y), v[k].z - tf32_trunc(v[k].z),
                                                                 v[k].w - tf32_trunc(v[k].w));
            }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&lo_ready[ls]);
                ptx::mbar_arrive(&hi_free[s]);
            }
        }
        tm.flush(st == 0);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

static int tc_planes(const FusedParams &p) { return p.window_stat == FUSED_STAT_MINMAX ? 2 : 1; }
// two lo tiles when they fit into the 227 KB a CTA can have
int tc_lo_stages(const FusedParams &p, int hp) {
    return 1024 + TcSmem::total(p.time_range * hp, p.n_out, tc_planes(p), 2) <= 227 * 1024 ? 2 : 1;
}
size_t tc_smem_bytes(const FusedParams &p, int hp) {
    return 1024 + TcSmem::total(p.time_range * hp, p.n_out, tc_planes(p), tc_lo_stages(p, hp));
}
int tc_tile_frames() { return kTileFrames; }
int tc_k_pad() { return kKPad; }
int tc_max_n0() { return kMaxN0; }
bool tc_layout_fits(int time_range, int n0) {
    // product ring: tile it (start b) reuses the positions of tile it - groups (start b + ring - groups * tile), whose last T-1
    // rows the evaluations of tile it - groups + 1 may still read: ring - groups * tile >= T - 1.
    // statistic ring: the MMA warp cannot pass the DFT of tile it before the copy of tile it-3 is done, the spectrum warps are
    // at most at tile it, the slowest group at tile it - 2 - groups: (groups + 3) * tile + T columns.
    return n0 <= kMaxN0 && kNumGroups * kTileFrames + time_range <= kPRing &&
           (kNumGroups + 3) * kTileFrames + time_range <= kStatRing;
}

cudaError_t launch_tc(int hp, int grid, size_t smem, const FusedParams &p, const TcWork &w, const void *tmap_main, const void *tmap_tail,
                      cudaStream_t stream) {
    const CUtensorMap &tm = *static_cast<const CUtensorMap *>(tmap_main);
    const CUtensorMap &tt = *static_cast<const CUtensorMap *>(tmap_tail);
    cudaError_t e = cudaErrorInvalidValue;
    auto go = [&](auto kern) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kern<<<grid, kTcThreads, smem, stream>>>(p, w, tm, tt);
    };
    const bool scaled = p.scaling != SYLDET_SCALING_LINEAR;
    const bool fast = hp == 4 && !scaled && p.window_stat == FUSED_STAT_L2 && p.tf[0] == SYLDET_TF_TANSIG && p.n_layers == 2 && p.n_out == 1 &&
                      p.tf[1] == SYLDET_TF_PURELIN && p.n_op == 1;
    const bool f16 = fast && w.f16_corr;   // the fp16 correction pass exists for the sample shape only
    if (w.debug_timing) {   // SYLDET_TC_TIMING: instrumented build of the common shape only
        if (f16) go(tc_detect_kernel<4, false, true, true, true>);
        else if (fast) go(tc_detect_kernel<4, false, true, true, false>);
        else return cudaErrorNotSupported;
    } else if (f16) go(tc_detect_kernel<4, false, false, true, true>);
    else if (fast) go(tc_detect_kernel<4, false, false, true, false>);
    else if (hp == 4 && !scaled) go(tc_detect_kernel<4, false, false, false, false>);
    else if (hp == 4) go(tc_detect_kernel<4, true, false, false, false>);
    else if (!scaled) go(tc_detect_kernel<8, false, false, false, false>);
    else go(tc_detect_kernel<8, true, false, false, false>);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace syldet
