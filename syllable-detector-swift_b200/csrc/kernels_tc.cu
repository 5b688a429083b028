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
// Roles (18 warps, one CTA per SM, persistent over (channel, chunk) units; every role walks the same tile sequence):
//   warp 0        TMA producer                      full[s] <- hi_free[s]
//   warp 1        MMA issuer + TMEM allocator       DFT(it): full, tmem_empty (even it; implied for odd it), lo_ready -> hi_free, tmem_full, lo_free
//                                                   layer0(pair k = tiles 2k, 2k+1), issued after DFT(2k+2): a_ready, p_empty -> p_full, a_free
//   then 4        evaluators (F), one warp per TMEM lane quadrant: layer 0 is ONE M = 128 contraction per PAIR of tiles (rows 0-63:
//                                                   tile 2k, rows 64-127: tile 2k+1), so every lane owns one product row = one evaluation:
//                                                   p_full -> product ring -> p_empty; diagonal sum (T x LDS.128), window statistic,
//                                                   network tail, events
//   then 8        spectrum warps (D)                tmem_full -> D -> registers -> tmem_empty; two shuffle rounds -> |X| ->
//                                                   layer-0 A operand (half of the pair's 128 rows) + per-column statistic partials -> a_ready
//   then 4        splitters (S)                     full, lo_free -> lo tile -> lo_ready, hi_free
// Every role is a serial chain per tile and runs at ~0.1 IPC per warp (profiles/): the roles overlap through double buffers.
// Round 1 ran layer 0 per tile (M = 64: half-rate tensor pipe, 16 of 32 lanes per evaluator warp, two evaluator groups of four
// warps taking tiles in turn); the pair scheme halves the evaluators' instruction count per evaluation and the weight-operand
// reads, and frees four warps (96 registers per thread). The direct variant (kDirect, opt-in: SYLDET_TC_DIRECT=1) replaces warp 0's TMA
// loads and the raw tile by six splitter warps that read the audio from global memory themselves (20 warps).
#include <cuda.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "fused_epilogue.cuh"
#include "ptx_sm100.cuh"

namespace syldet {

namespace {

constexpr int kFGroup = 4;                     // evaluator warps (one per TMEM lane quadrant)
// kF16: the whole band DFT as kind::f16 MMAs on two-term fp16 splits of both operands (linear scaling + a per-window normaliser):
//   A = a1 + a2, a1 = fp16(A), a2 = fp16(A - a1);   x = h1 + h2 * 2^-11, h1 = fp16(x), h2 = fp16((x - h1) * 2^11)
//   A x ~= a1 h1 + (a1 2^-11) h2 + a2 h1            (the dropped term a2 (x - h1) is <= 2^-24 |A| |x|)
// as ONE K-concatenated pass [a1 | a1 2^-11 | a2] * [h1 ; h2 ; h1]: 26 MMAs with K = 16 per tile, where round 1 ran 17 kind::tf32
// MMAs on the raw audio plus 17 kind::f16 correction MMAs (a tf32 MMA takes as long as an f16 one and moves as many operand bytes
// for half the K). The splitters write the fp16 tile with the byte geometry of the fp32 one (4 x [64 rows x 128 B] + [64 rows x 32 B]):
// chunks 0,1 = h1(n < 128), chunks 2,3 = h2(n < 128), tail = h1(n >= 128) | h2(n >= 128); the third product re-reads chunks 0,1 and
// the tail (against a zero A block for its h2 half). The raw fp32 tile is then read by the splitters only.
// fp16's 11 bits per term keep the sum at float32 level for |x| in [6e-5, 32752]; quieter samples carry an absolute error of
// ~2^-36 each, which the evaluators' range guard bounds against the window's energy.
#ifndef TC_NUM_D
#define TC_NUM_D 8
#endif
constexpr int kWarpTma = 0, kWarpMma = 1, kWarpF0 = 2, kNumF = kFGroup, kWarpD0 = kWarpF0 + kNumF, kNumD = TC_NUM_D,
              kWarpS0 = kWarpD0 + kNumD, kNumS = 4;
constexpr int kNumSDirect = 6;                 // splitter warps of the direct variant (they also carry the global loads)
constexpr int kTcThreadsDirect = (kWarpS0 + kNumSDirect) * 32;  // 640 (20 warps): still 96 registers per thread
constexpr int kTcThreads = (kWarpS0 + kNumS) * 32;  // 576 (18 warps) with 8 spectrum warps; 832 with -DTC_NUM_D=16 (slower: profiles/r02_tc_v21_numd16_vs_8.txt)
static_assert(kNumD == 8 || kNumD == 16, "spectrum warps: two or four per TMEM lane quadrant");
static_assert(kWarpF0 % 4 == 2 && kWarpD0 % 4 == 2, "quadrant / half assignment below assumes these starts");
constexpr int kTileRows = 64;                  // rows of Y per tile = N of the DFT MMA
constexpr int kDSplit = kNumD / 4;             // spectrum warps per TMEM lane quadrant
constexpr int kDCols = kTileRows / kDSplit;    // tile columns per spectrum warp (32 or 16)
constexpr int kDGroups = kDCols / 4;           // groups of four columns
constexpr int kDMags = kDCols / 4;             // magnitudes a lane ends up with
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
// kF16: the A operand [a1(128) | a1 2^-11 (128) | a1(8) a1 2^-11 (8) | a2(128) | a2(8) 0(8)] in fp16 pairs, from kColAhi
constexpr int kA16Cols = 64 + 64 + 8 + 64 + 8;
static_assert(kA16Cols <= 2 * kKPad, "the fp16 A operand fits where the tf32 hi / lo parts go");
static_assert(kColP0 + 2 * kMaxN0 <= kTmemCols, "TMEM budget");
constexpr int kPairFrames = 2 * kTileFrames;   // evaluations the evaluators finish per pass (one pair of tiles)
constexpr int kPRing = kPairFrames + 22;       // product-row ring (rows = columns): the pair in flight + the T-1 (<= 22) rows before it
constexpr int kStatRing = 640;                 // per-column statistic ring: [2 planes][kStatRing][4 bin quarters] float (see tc_layout_fits)
constexpr int kBarF = 2;                       // named barrier of the evaluators
constexpr int kEvCap = 192;                    // shared-memory event buffer of the evaluators (flushed with one global atomic)
static_assert(kEvCap >= kPairFrames + 32, "the event buffer must take one pair after every flush check");

struct TcSmem {  // byte offsets from the 1024-byte aligned base
    static constexpr int hi0 = 0, hi1 = kTileBytes, lo = 2 * kTileBytes;   // audio tiles (34 816 B = 34 swizzle atoms each)
    static constexpr int abuf = 3 * kTileBytes;                     // [hi, lo][128 rows x 128 B]: rows 0-63 even tile, 64-127 odd tile of a pair
    static constexpr int wcat = abuf + 4 * 8192;                    // [hi, lo][<= 56 rows x 128 B]
    static constexpr int pbuf = wcat + 2 * kMaxN0 * 128;            // [kPRing][ppitch] float; everything after it is placed at run time
    // then: float4 colstat[planes][kStatRing] | event meta int4[kEvCap] | event outputs float[kEvCap][n_out] |
    //       barriers (256 B) | optionally the second lo tile (1024-byte aligned) when it fits
    // planes: 1 (sum of squares), 2 for the min/max statistic
    __host__ __device__ static constexpr int ppitch(int np) { return ((((np + 7) >> 3) << 1) | 1) << 2; }  // whole 8-float chunks + 1: an odd number of float4
    __host__ __device__ static constexpr int colstat(int np) { return pbuf + kPRing * ppitch(np) * 4; }
    __host__ __device__ static constexpr int evmeta(int np, int planes) { return colstat(np) + planes * kStatRing * 16; }
    __host__ __device__ static constexpr int evout(int np, int planes) { return evmeta(np, planes) + kEvCap * 16; }
    __host__ __device__ static constexpr int bars(int np, int n_out, int planes) { return evout(np, planes) + kEvCap * n_out * 4; }
    __host__ __device__ static constexpr int lo1(int np, int n_out, int planes) { return (bars(np, n_out, planes) + 256 + 1023) & ~1023; }
    __host__ __device__ static constexpr int total(int np, int n_out, int planes, int lo_stages) {
        return lo_stages == 2 ? lo1(np, n_out, planes) + kTileBytes : bars(np, n_out, planes) + 256;
    }
    __host__ __device__ static constexpr int hi(int stage) { return stage ? hi1 : hi0; }
    __host__ __device__ static constexpr int a(int half, int part) { return abuf + part * 16384 + half * 8192; }
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

// Two-term fp16 split of four samples: hx = fp16(x) (pairs), hl = fp16((x - hx) * 2^11). sm_100 has mixed-precision FMAs
// (fma.rn.f32.f16: f16 x f16 + f32, SASS FHFMA with .H0/.H1 selectors), so the exact residual times 2^11 is ONE instruction per
// sample on top of x * 2^11 (FMUL2, two samples per instruction): 10 instructions per float4 where unpack / subtract / multiply took 16.
__device__ __forceinline__ void split_fp16(const float4 &v, uint2 &hx, uint2 &hl) {
    const __half2 x01 = __floats2half2_rn(v.x, v.y), x23 = __floats2half2_rn(v.z, v.w);
    hx = make_uint2(*reinterpret_cast<const uint32_t *>(&x01), *reinterpret_cast<const uint32_t *>(&x23));
    float r0, r1, r2, r3;
    const unsigned short m2048 = 0xE800;   // -2048 in fp16
    const float2 k2048 = make_float2(2048.0f, 2048.0f);
    const float2 t01 = ptx::mul2(make_float2(v.x, v.y), k2048), t23 = ptx::mul2(make_float2(v.z, v.w), k2048);   // FMUL2
    asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nfma.rn.f32.f16 %0, lo, %5, %3;\nfma.rn.f32.f16 %1, hi, %5, %4;\n}\n"
        : "=f"(r0), "=f"(r1)
        : "r"(hx.x), "f"(t01.x), "f"(t01.y), "h"(m2048));
    asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nfma.rn.f32.f16 %0, lo, %5, %3;\nfma.rn.f32.f16 %1, hi, %5, %4;\n}\n"
        : "=f"(r2), "=f"(r3)
        : "r"(hx.y), "f"(t23.x), "f"(t23.y), "h"(m2048));
    const __half2 l01 = __floats2half2_rn(r0, r1), l23 = __floats2half2_rn(r2, r3);
    hl = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
}

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
template <int HP, bool kScaled, bool kTiming, bool kFast, bool kF16, bool kDirect>
__global__ void __launch_bounds__(kDirect ? kTcThreadsDirect : kTcThreads, 1)
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
    uint64_t *full = bars, *hi_free = bars + 2, *lo_ready = bars + 4, *tmem_full = bars + 6, *tmem_empty = bars + 8;   // [2] each
    uint64_t *p_full = bars + 10, *p_empty = bars + 12, *lo_free = bars + 14;                                           // [2] each
    uint64_t *a_ready = bars + 16, *a_free = bars + 17;                       // one A operand (128 rows), refilled pair by pair
    // kDirect: no TMA, no raw tile. The splitters read the audio from global memory and write the fp16 tile straight from
    // registers; the fp16 tiles cycle through n_stages = 3 or 4 buffers (the two raw-tile buffers, lo, and lo1 when it fits).
    static_assert(!kDirect || kF16, "the direct variant feeds the fp16 band DFT");
    if constexpr (kDirect) {
        lo_ready = bars + 22;     // [4]
        lo_free = bars + 26;      // [4]
    }
    const int n_stages = w.lo_stages == 2 ? 4 : 3;
    // lo tile(s): with two, stage = it & 1 like the hi tiles and the splitters never wait for pass 3 of the previous tile
    const int lo_stages = w.lo_stages;
    const int lo1_off = TcSmem::lo1(np, p.n_out, planes);
    auto stage_off = [&](int i) { return i < 3 ? i * kTileBytes : lo1_off; };   // kDirect: fp16 tile buffers
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 18);
    int *ev_count = reinterpret_cast<int *>(bars + 19);                       // events waiting in shared memory
    unsigned long long *ev_base = reinterpret_cast<unsigned long long *>(bars + 20);
    float *pbuf = reinterpret_cast<float *>(smem + TcSmem::pbuf);
    float *colstat = reinterpret_cast<float *>(smem + TcSmem::colstat(np));   // sum of squares | minimum, then the maximum plane

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kThreads = kDirect ? kTcThreadsDirect : kTcThreads;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&hi_free[i], kF16 ? kNumS : 1 + kNumS);   // kF16: the tensor core never reads the raw tile
            ptx::mbar_init(&tmem_full[i], 1);
            ptx::mbar_init(&tmem_empty[i], kNumD);
            ptx::mbar_init(&p_full[i], 1);
            ptx::mbar_init(&p_empty[i], kFGroup);
            ptx::mbar_init(&lo_ready[i], kNumS);
            ptx::mbar_init(&lo_free[i], 1);
        }
        if constexpr (kDirect) {
            for (int i = 0; i < 4; ++i) {
                ptx::mbar_init(&lo_ready[i], kNumSDirect);
                ptx::mbar_init(&lo_free[i], 1);
            }
        }
        ptx::mbar_init(a_ready, 2 * kNumD);   // both tiles of a pair
        ptx::mbar_init(a_free, 1);
        *ev_count = 0;
        ptx::fence_mbar_init();
    }
    if (warp == kWarpMma) {
        ptx::tmem_alloc(tmem_ptr, kTmemCols);
        ptx::tmem_relinquish();
    }
    // layer-0 weights -> shared memory, K-major SWIZZLE_128B rows of 32 bins (B operand of the second contraction)
    for (int i = tid; i < 2 * n0 * 32; i += kThreads) {
        const int part = i / (n0 * 32), r = (i / 32) % n0, c = i % 32;
        *reinterpret_cast<float *>(smem + TcSmem::wcat + part * kMaxN0 * 128 + sw128(r, c)) = __ldg((part ? w.wcat_lo : w.wcat_hi) + r * 32 + c);
    }
    // rows of the layer-0 A operand that no tile writes (row 63, short last tiles) must hold finite values
    for (int i = tid; i < 4 * 8192 / 16; i += kThreads) reinterpret_cast<float4 *>(smem + TcSmem::abuf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (kDirect) {   // the K padding of the fp16 tiles (n >= hop) is never written: it must read as zero
        for (int i = tid; i < 3 * kTileBytes / 16; i += kThreads) reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n_stages == 4)
            for (int i = tid; i < kTileBytes / 16; i += kThreads) reinterpret_cast<float4 *>(smem + lo1_off)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
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
        if constexpr (kF16) {
            const uint32_t *src = w.dft16 + (size_t)m * kA16Cols;
            for (int kb = 0; kb < kA16Cols / 8; ++kb) {
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = __ldg(src + kb * 8 + i);
                ptx::tmem_st_x8(lane_addr + kColAhi + kb * 8, r);
            }
        } else {
            for (int part = 0; part < 2; ++part) {
                const float *src = (part ? w.dft_lo : w.dft_hi) + (size_t)m * kKPad;
                for (int kb = 0; kb < kKPad / 8; ++kb) {
                    uint32_t r[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__ldg(src + kb * 8 + i));
                    ptx::tmem_st_x8(lane_addr + (part ? kColAlo : kColAhi) + kb * 8, r);
                }
            }
        }
        ptx::tc_wait_st();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    if (warp == kWarpTma) {
        // ================================ TMA producer ================================================================
        if constexpr (kDirect) {
            // L2 prefetcher of the direct variant: one bulk prefetch per tile, w.pf_dist tiles ahead of the tile the splitters are
            // about to fill (paced by the same lo_free barriers they wait on)
            if (w.pf_dist > 0 && ptx::elect_one()) {
                TileWalk tw, pf;
                tw.init(w, T);
                pf = tw;
                auto prefetch = [&]() {
                    if (!pf.valid()) return;
                    const int rows = min(kTileRows, w.n_rows - pf.first_row());
                    if (rows > 0)
                        if (!w.pcm16)   // (16-bit sources: no prefetch)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(w.pcm + (int64_t)pf.ch * w.ch_stride + (int64_t)pf.first_row() * p.hop),
                                     "r"(rows * p.hop * 4)
                                     : "memory");
                    pf.next(w, T);
                };
                for (int i = 0; i < w.pf_dist; ++i) prefetch();
                int stg = 0;
                uint32_t stg_use = 0;
                for (; tw.valid(); tw.next(w, T)) {
                    ptx::mbar_wait_relaxed(&lo_free[stg], (stg_use & 1) ^ 1);
                    if (++stg == n_stages) { stg = 0; ++stg_use; }
                    prefetch();
                }
            }
        } else if (ptx::elect_one()) {
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
            // kF16: the fp16 band DFT, same descriptors and column steps (32 B of K per instruction = 16 k = 8 TMEM columns).
            // Segments of the A operand (TMEM columns from kColAhi): [0, 128) a1 | a1 2^-11 against chunks 0..3 (h1, h2), [128, 136)
            // against the tail, [136, 200) a2 against chunks 0,1 (h1 again), [200, 208) a2 | 0 against the tail.
            constexpr uint32_t idesc_f16 = ptx::idesc_f16(128, kTileRows);
            auto f16_pass = [&](uint32_t d, uint32_t a, uint32_t b) {
                const uint64_t desc0 = ptx::smem_desc_kmajor(b, 1024, 2), desc_tail = ptx::smem_desc_kmajor(b + kMainChunks * kMainBytes, 256, 6);
                uint32_t acc = 0;
#pragma unroll 1
                for (int j = 0; j < kMainChunks + 2; ++j) {        // chunks 0, 1, 2, 3, then 0, 1 again
                    const int chunk = j < kMainChunks ? j : j - kMainChunks;
                    const uint32_t aj = a + (j < kMainChunks ? j * 32 : 136 + (j - kMainChunks) * 32);
                    const uint64_t desc = desc0 + (uint64_t)(chunk * (kMainBytes >> 4));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        ptx::mma_f16_ts(d, aj + ks * 8, desc + ks * 2, idesc_f16, acc);
                        acc = 1;
                    }
                }
                ptx::mma_f16_ts(d, a + 128, desc_tail, idesc_f16, 1);
                ptx::mma_f16_ts(d, a + 200, desc_tail, idesc_f16, 1);
            };
            RoleTimer<kTiming> tm(w.debug_timing, 6, false);   // the MMA issuer is the pacemaker: it polls
            const uint32_t idesc_l0 = ptx::idesc_tf32(128, n0);
            const uint32_t a_hi = ptx::smem_addr(smem + TcSmem::a(0, 0)), a_lo = ptx::smem_addr(smem + TcSmem::a(0, 1));
            // The issuer never blocks on one barrier while other work is ready (round 1 and the first pair-scheme build waited for
            // the pair's magnitudes before they queued the next tile's DFT, which left the tensor pipe idle ~28 % of a tile): it
            // probes what each of its two jobs needs (mbarrier.test_wait) and issues whichever is ready - the band DFT of the next
            // tile (its fp16 / lo tile written, its accumulator read by the spectrum warps two tiles ago) or layer 0 of the next
            // pair (both tiles' magnitudes written, the product buffer drained) - so the DFT runs up to two tiles ahead.
            int stg = 0;          // kDirect: fp16 tile buffer of tile `it` and how often it has been filled before
            uint32_t stg_use = 0;
            auto dft_ready = [&](uint32_t i) {
                const int s = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                if constexpr (kDirect) return ptx::mbar_test(&tmem_empty[s], ph ^ 1) && ptx::mbar_test(&lo_ready[stg], stg_use & 1);
                const int ls = lo_stages == 2 ? s : 0;
                const uint32_t lo_use = lo_stages == 2 ? (i >> 1) : i;   // uses of this lo buffer so far
                if constexpr (!kF16) {
                    if (!ptx::mbar_test(&full[s], ph)) return false;      // hi landed (TMA); kF16 reads only what the splitters wrote
                }
                return ptx::mbar_test(&tmem_empty[s], ph ^ 1) && ptx::mbar_test(&lo_ready[ls], lo_use & 1);
            };
            auto issue_dft = [&](uint32_t i) {
                const int s = i & 1;
                const int ls = lo_stages == 2 ? s : 0;
                ptx::tc_fence_after();
                const uint32_t d = tmem_base + kColD0 + s * kTileRows;
                const uint32_t hi = ptx::smem_addr(smem + TcSmem::hi(s));
                if constexpr (kDirect) {
                    f16_pass(d, tmem_base + kColAhi, ptx::smem_addr(smem + stage_off(stg)));
                    ptx::mma_commit(&tmem_full[s]);
                    ptx::mma_commit(&lo_free[stg]);
                    if (++stg == n_stages) { stg = 0; ++stg_use; }
                    return;
                } else if constexpr (kF16) {
                    f16_pass(d, tmem_base + kColAhi, ls ? lo_b : lo_a);
                } else {
                    dft_pass(d, tmem_base + kColAhi, hi, 0);
                    dft_pass(d, tmem_base + kColAlo, hi, 1);
                    ptx::mma_commit(&hi_free[s]);         // the MMA side is done with hi[s]
                    dft_pass(d, tmem_base + kColAhi, ls ? lo_b : lo_a, 1);
                }
                ptx::mma_commit(&tmem_full[s]);
                ptx::mma_commit(&lo_free[ls]);
            };
            auto l0_ready = [&](uint32_t pk) {
                return ptx::mbar_test(a_ready, pk & 1) && ptx::mbar_test(&p_empty[pk & 1], ((pk >> 1) & 1) ^ 1);
            };
            auto issue_l0 = [&](uint32_t pk) {  // per-column layer-0 products of the pair pk = tiles 2pk, 2pk+1 (M = 128)
                const int pb = pk & 1;
                ptx::tc_fence_after();
                const uint32_t d = tmem_base + kColP0 + pb * kMaxN0;
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
                ptx::mma_commit(&p_full[pb]);
                ptx::mma_commit(a_free);
            };
            TileWalk tw;
            tw.init(w, T);
            uint32_t it = 0, next_pair = 0;   // next tile whose DFT is to be issued; next pair whose layer 0 is to be issued
            bool more = tw.valid();
            while (more || next_pair < (it + 1) / 2) {
                bool progress = false;
                if (more && dft_ready(it)) {
                    issue_dft(it);
                    ++it;
                    tw.next(w, T);
                    more = tw.valid();
                    progress = true;
                }
                const uint32_t pairs_queued = more ? it / 2 : (it + 1) / 2;   // pairs whose tiles all have their DFT issued
                if (next_pair < pairs_queued && l0_ready(next_pair)) {
                    issue_l0(next_pair++);
                    progress = true;
                }
                if (!progress) {
                    const long long t0 = tm.now();
                    __nanosleep(20);
                    tm.add(3, t0);   // idle: nothing ready
                }
            }
            tm.flush(true);
        }
    } else if (warp < kWarpD0) {
        // ================================ evaluators (F) ==============================================================
        // Four warps, one per TMEM lane quadrant. The layer-0 accumulator P[pk & 1] of pair pk is an M = 128 tile: product row r sits
        // in lane r; rows 0-63 are the columns of tile 2pk, rows 64-127 those of tile 2pk+1. So quadrants 0/1 own tile 2pk (columns
        // 0-31 / 32-63) and quadrants 2/3 tile 2pk+1, and every lane owns one row: it moves the row to the product ring and then
        // evaluates the network whose newest column is that row.
        const int quad = warp & 3;
        const int ft = (warp - kWarpF0) * 32 + lane;        // thread index inside the role
        const int sel = quad >> 1;                          // which tile of the pair
        const int c = (quad & 1) * 32 + lane;               // column of that tile this thread owns
        int4 *ev_meta = reinterpret_cast<int4 *>(smem + TcSmem::evmeta(np, planes));   // (channel, -, eval lo, eval hi)
        float *ev_out = reinterpret_cast<float *>(smem + TcSmem::evout(np, planes));
        const int nchunks = (np + 7) >> 3;                  // 8-column chunks of a product row
        TileWalk tw;
        tw.init(w, T);
        int gcol = 0;                                       // product-ring position of the pair's first column (mod kPRing)
        int scol = 0;                                       // statistic-ring position of the same column (mod kStatRing)
        RoleTimer<kTiming> tm(w.debug_timing, 12);
        auto flush_events = [&](int n_ev) { flush_group_events(w.sink, ev_meta, ev_out, ev_count, ev_base, n_ev, ft, kBarF, n_out); };
        for (uint32_t pk = 0; tw.valid(); ++pk) {
            // this thread's tile of the pair: unit position, frame count, ring positions
            TileWalk mine = tw;
            const int frames_a = tw.frames();
            tw.next(w, T);
            int frames_b = 0;
            if (tw.valid()) {
                frames_b = tw.frames();
                if (sel) mine = tw;
                tw.next(w, T);
            }
            const int frames = sel ? frames_b : frames_a;   // 0: the last pair has no second tile
            int g_mine = gcol + (sel ? frames_a : 0), s_mine = scol + (sel ? frames_a : 0);
            if (g_mine >= kPRing) g_mine -= kPRing;
            if (s_mine >= kStatRing) s_mine -= kStatRing;
            const int pb = pk & 1;
            tm.wait(&p_full[pb], (pk >> 1) & 1, 0);
            ptx::tc_fence_after();
            const long long t_f0 = tm.now();
            {   // P row -> product ring, four 8-column chunks at a time: loads in flight, one wait, then the stores
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + kColP0 + pb * kMaxN0;
                int row = g_mine + c;
                if (row >= kPRing) row -= kPRing;
                float4 *dst = reinterpret_cast<float4 *>(pbuf + row * ppitch);
                const bool store = c < frames;
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
                if (lane == 0) ptx::mbar_arrive(&p_empty[pb]);
            }
            tm.sync(kBarF, kFGroup * 32, 1);   // the pair's rows (and, from the previous pass, the T-1 rows before them) are in the ring
            const long long t_f1 = tm.now();
            // evaluation whose newest column is column c of this thread's tile: unit-local index j
            const int j = mine.cols_before() - (T - 1) + c;
#ifdef TC_EXP_SKIP_EVAL
            const bool valid = false;
#else
            const bool valid = c < frames && j >= 0;
#endif
            bool hit = false;
            float out[kFusedMaxOut];
            {
                float acc[HP];
                float2 acc2[HP / 2];            // the same sums as register pairs (FADD2: two per issue slot)
#pragma unroll
                for (int h = 0; h < HP / 2; ++h) acc2[h] = make_float2(0.0f, 0.0f);
                float2 s2 = make_float2(0.0f, 0.0f);
                float s0 = window_stat == FUSED_STAT_L2 ? 0.0f : INFINITY, s1 = -INFINITY;
                if (valid) {
                    int row = g_mine + c - (T - 1);                        // ring position of the window's oldest column
                    if (row < 0) row += kPRing;
                    else if (row >= kPRing) row -= kPRing;
                    int col = s_mine + c - (T - 1);                        // same column in the statistic ring
                    if (col < 0) col += kStatRing;
                    else if (col >= kStatRing) col -= kStatRing;
                    const float *pt = pbuf + row * ppitch;                 // advances by one ring row and one (t, :) block per step
                    const float4 *cs = reinterpret_cast<const float4 *>(colstat) + col;
#pragma unroll 2
                    for (int t = 0; t < T; ++t) {
                        const float4 *prow = reinterpret_cast<const float4 *>(pt);
                        pt += ppitch + HP;
                        if (++row == kPRing) { row = 0; pt -= kPRing * ppitch; }
                        const float4 v0 = prow[0];
                        acc2[0] = ptx::add2(acc2[0], make_float2(v0.x, v0.y));
                        acc2[1] = ptx::add2(acc2[1], make_float2(v0.z, v0.w));
                        if constexpr (HP == 8) {
                            const float4 v1 = prow[1];
                            acc2[2] = ptx::add2(acc2[2], make_float2(v1.x, v1.y));
                            acc2[3] = ptx::add2(acc2[3], make_float2(v1.z, v1.w));
                        }
                        const float4 ca = cs[0];            // the four bin quarters of the column
                        if (window_stat == FUSED_STAT_L2) s2 = ptx::add2(s2, ptx::add2(make_float2(ca.x, ca.y), make_float2(ca.z, ca.w)));
                        else if (window_stat == FUSED_STAT_MINMAX) {
                            const float4 cb = cs[kStatRing];
                            s0 = fminf(s0, fminf(fminf(ca.x, ca.y), fminf(ca.z, ca.w)));
                            s1 = fmaxf(s1, fmaxf(fmaxf(cb.x, cb.y), fmaxf(cb.z, cb.w)));
                        }
                        ++cs;
                        if (++col == kStatRing) { col = 0; cs -= kStatRing; }
                    }
                }
#pragma unroll
                for (int h = 0; h < HP / 2; ++h) {
                    acc[2 * h] = acc2[h].x;
                    acc[2 * h + 1] = acc2[h].y;
                }
                if (window_stat == FUSED_STAT_L2) s0 = s2.x + s2.y;
                if constexpr (kF16) {
                    // Range guard of the fp16 correction pass: the pass is at float32 level while the window's band
                    // energy sum |X|^2 is at least guard_lo (absolute operand errors of 2^-25 stay below 1e-6 of the normalised
                    // features) and no sample overflowed fp16 (inf / NaN energy otherwise). An exactly silent window (energy 0) is
                    // exact in both variants. Anything else raises the flag; the host then repeats the launch with the all-TF32 variant.
                    // Windows normalised by their minimum / maximum (mapminmax over the window): the same argument with the
                    // window's range max - min in the place of its norm (a flat window is exact: every input becomes -1).
                    if (window_stat == FUSED_STAT_L2) {
                        if (valid && !(s0 >= w.guard_lo && s0 <= 3.0e38f) && s0 != 0.0f) *w.range_flag = 1;
                    } else if (window_stat == FUSED_STAT_MINMAX) {
                        const float range = s1 - s0;
                        if (valid && !(range >= w.guard_range && s1 <= 3.0e38f) && range != 0.0f) *w.range_flag = 1;
                    }
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
                float *o = w.all_out + ((int64_t)mine.ch * w.out_evals_per_channel + w.eval_offset + mine.e0 + j) * n_out;
                const bool store = valid && w.all_out != nullptr;
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
                hit = hit && valid;
            }
            const unsigned hits = __ballot_sync(0xffffffffu, hit);
            if (hits) {   // append to the shared-memory event buffer (room for a whole pair is guaranteed by the flush rule)
                int base = 0;
                if (lane == 0) base = atomicAdd(ev_count, __popc(hits));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) {
                    const int e = base + __popc(hits & ((1u << lane) - 1));
                    const int64_t ev = w.eval_offset + mine.e0 + j;
                    ev_meta[e] = make_int4(mine.ch, 0, (int)(unsigned)(ev & 0xffffffffll), (int)(ev >> 32));
#pragma unroll 1
                    for (int k = 0; k < n_out; ++k) ev_out[e * n_out + k] = pick(out, k);
                }
            }
            tm.sync(kBarF, kFGroup * 32, 2);  // the next pair overwrites ring rows that this pair's evaluations read
            const int n_ev = *ev_count;
            if (n_ev > kEvCap - kPairFrames) flush_events(n_ev);
            gcol += frames_a + frames_b;
            if (gcol >= kPRing) gcol -= kPRing;
            scol += frames_a + frames_b;
            if (scol >= kStatRing) scol -= kStatRing;
        }
        {
            const int n_ev = *ev_count;
            if (n_ev > 0) flush_events(n_ev);
        }
        tm.flush(ft == 0);
    } else if (warp < kWarpS0) {
        // ================================ spectrum warps (D) ===========================================================
        // TMEM lane 32*quad + lane holds DFT-matrix row (part = lane >> 3: Re B1 | Im B1 | Re B2 | Im B2, bin = 8*quad + (lane & 7)),
        // so the four parts of a bin sit in one warp and combine through two shuffle rounds; nothing goes through shared memory:
        //   (1) X[c] = P1[c] + P2[c+1]: B1 lanes (bit 4 clear) take columns 4g, 4g+1, B2 lanes take 4g+2, 4g+3 (xor 16)
        //   (2) re^2 + im^2: Re lanes (bit 3 clear) keep the columns of even g, Im lanes those of odd g           (xor 8)
        // after which every lane owns kDMags magnitudes of its bin: columns col0 + 8*i + u. A store instruction then covers rows
        // r, r+2, r+4, r+6 x 8 bins = all 32 banks of the SWIZZLE_128B operand.
        // kDSplit warps share a quadrant, each taking kDCols consecutive columns of the tile: this role is a dependent chain of TMEM
        // load -> shuffles -> MUFU -> shuffles -> stores, and with the evaluators out of the way it is what the MMA issuer waits for;
        // four warps per quadrant halve that chain (round 1 and the first pair-scheme build ran two).
        const int quad = warp & 3, dw = warp - kWarpD0, part = dw >> 2;
        const bool up = (lane & 16) != 0, im = (lane & 8) != 0;
        const int bin = quad * 8 + (lane & 7);
        const bool in_band = bin < L;
        const int col0 = part * kDCols + (up ? 2 : 0) + (im ? 4 : 0);
        int a_off[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) a_off[u] = (col0 + u) * 128 + ((((bin >> 2) ^ (col0 + u)) & 7) << 4) + ((bin & 3) << 2);
        // the column whose statistic ends up in this lane: value index e = 2*i + u is selected by lane bits 2.. (see reduce below)
        const int stat_col = kDMags == 8 ? col0 + 8 * ((lane & 7) >> 1) + (lane & 1) : col0 + 8 * ((lane >> 2) & 1) + ((lane >> 1) & 1);
        const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + kColD0 + part * kDCols;
        TileWalk tw;
        tw.init(w, T);
        int gcol = 0;                                       // statistic-ring position of the tile's first column (mod kStatRing)
        RoleTimer<kTiming> tm(w.debug_timing, 18);
        uint32_t it = 0;
        for (; tw.valid(); ++it, tw.next(w, T)) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int frames = tw.frames();
            tm.wait(&tmem_full[s], ph, 0);
            ptx::tc_fence_after();
            // One load of the warp's columns (+1), then straight-line code: the shuffle chains of the column groups are
            // independent, and a warp needs that many in flight to hide the latency of the (busy) shared-memory/shuffle pipe.
            uint32_t r[kDCols + 1];
            const long long t_d0 = tm.now();
            {
                r[kDCols] = 0;                                  // column 64 does not exist: frame 63 is never complete in this tile
                if constexpr (kDCols == 32) {
                    uint32_t r32[32];
                    ptx::tmem_ld_x32(taddr0 + s * kTileRows, r32);
                    if (part + 1 < kDSplit) ptx::tmem_ld_x1(taddr0 + s * kTileRows + kDCols, r[kDCols]);
                    ptx::tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = r32[j];
                } else {
                    uint32_t r16[16];
                    ptx::tmem_ld_x16(taddr0 + s * kTileRows, r16);
                    if (part + 1 < kDSplit) ptx::tmem_ld_x1(taddr0 + s * kTileRows + kDCols, r[kDCols]);
                    ptx::tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = r16[j];
                }
            }
            tm.add(4, t_d0);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[s]);
            // ---- (1) X_c = (Re1[c] + Re2[c+1]) + i (Im1[c] + Im2[c+1]) ------------------------------------------------------
            float x[kDGroups][2];
#pragma unroll
            for (int g = 0; g < kDGroups; ++g) {
                float own[2], got[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float send = __uint_as_float(up ? r[4 * g + u + 1] : r[4 * g + 2 + u]);
                    own[u] = __uint_as_float(up ? r[4 * g + 3 + u] : r[4 * g + u]);
                    got[u] = __shfl_xor_sync(0xffffffffu, send, 16);
                }
                const float2 sum = ptx::add2(make_float2(own[0], own[1]), make_float2(got[0], got[1]));   // both columns in one FADD2
                x[g][0] = sum.x;
                x[g][1] = sum.y;
            }
            // ---- (2) |X| of the band (plain multiplies and add, as the reference computes it) -------------------------------
            float mag[kDMags];                                  // e = 2*i + u: column col0 + 8*i + u
#pragma unroll
            for (int i = 0; i < kDGroups / 2; ++i) {
                const float2 xa = make_float2(x[2 * i][0], x[2 * i][1]), xb = make_float2(x[2 * i + 1][0], x[2 * i + 1][1]);
                const float2 sqa2 = ptx::mul2(xa, xa), sqb2 = ptx::mul2(xb, xb);   // FMUL2: the squares of both columns
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float sq_a = u ? sqa2.y : sqa2.x, sq_b = u ? sqb2.y : sqb2.x;
                    const float other = __shfl_xor_sync(0xffffffffu, im ? sq_a : sq_b, 8);
                    float v;
                    if constexpr (kScaled) {
                        const float raw = sqrt_fast(__fadd_rn(im ? sq_b : sq_a, other));
                        if (w.debug_band && in_band && col0 + 8 * i + u < frames)   // extractPower() values, before the scaling
                            w.debug_band[((int64_t)tw.ch * w.debug_cols + w.eval_offset + tw.e0 + tw.cols_before() + col0 + 8 * i + u) * L + bin] = raw;
                        v = scale_value_nl(raw, p.scaling);
                    } else v = sqrt_fast_ftz(__fadd_rn(im ? sq_b : sq_a, other));
                    mag[2 * i + u] = in_band ? v : 0.0f;
                }
            }
            // ---- window statistic: per-column partial over this warp's 8 bins; the evaluators combine the four quadrants -----
            if (window_stat != FUSED_STAT_NONE) {
                // kDMags values x eight lanes -> the lane whose bits select value e holds its total: exchange half of the values per round
                auto reduce = [&](float (&q)[kDMags], auto op) {
                    const bool b4 = (lane & 4) != 0, b2 = (lane & 2) != 0, b1 = (lane & 1) != 0;
                    if constexpr (kDMags == 8) {
                        float h4[4], h2[2];
#pragma unroll
                        for (int k = 0; k < 4; ++k) h4[k] = op(b4 ? q[k + 4] : q[k], __shfl_xor_sync(0xffffffffu, b4 ? q[k] : q[k + 4], 4));
#pragma unroll
                        for (int k = 0; k < 2; ++k) h2[k] = op(b2 ? h4[k + 2] : h4[k], __shfl_xor_sync(0xffffffffu, b2 ? h4[k] : h4[k + 2], 2));
                        return op(b1 ? h2[1] : h2[0], __shfl_xor_sync(0xffffffffu, b1 ? h2[0] : h2[1], 1));
                    } else {
                        float h2[2];
#pragma unroll
                        for (int k = 0; k < 2; ++k) h2[k] = op(b4 ? q[k + 2] : q[k], __shfl_xor_sync(0xffffffffu, b4 ? q[k] : q[k + 2], 4));
                        const float h1 = op(b2 ? h2[1] : h2[0], __shfl_xor_sync(0xffffffffu, b2 ? h2[0] : h2[1], 2));
                        return op(h1, __shfl_xor_sync(0xffffffffu, h1, 1));   // both lanes of a pair end with the total
                    }
                };
                float q[kDMags];
                int sidx = gcol + stat_col;
                if (sidx >= kStatRing) sidx -= kStatRing;
                float *cs = colstat + (sidx << 2) + quad;
                const bool writer = (kDMags == 8 || (lane & 1) == 0) && stat_col < frames;
                if (window_stat == FUSED_STAT_L2) {
#pragma unroll
                    for (int e = 0; e < kDMags; e += 2) {
                        const float2 m2 = make_float2(mag[e], mag[e + 1]), sq = ptx::mul2(m2, m2);
                        q[e] = sq.x;
                        q[e + 1] = sq.y;
                    }
                    const float tot = reduce(q, [](float a, float b) { return a + b; });
                    if (writer) cs[0] = tot;
                } else {
#pragma unroll
                    for (int e = 0; e < kDMags; ++e) q[e] = in_band ? mag[e] : INFINITY;
                    const float mn = reduce(q, [](float a, float b) { return fminf(a, b); });
#pragma unroll
                    for (int e = 0; e < kDMags; ++e) q[e] = in_band ? mag[e] : -INFINITY;
                    const float mx = reduce(q, [](float a, float b) { return fmaxf(a, b); });
                    if (writer) {
                        cs[0] = mn;
                        cs[kStatRing * 4] = mx;
                    }
                }
            }
            // ---- magnitudes -> layer-0 A operand (hi, lo) -----------------------------------------------------------------------
            tm.wait(a_free, ph ^ 1, 2);                     // layer 0 of the previous pair has read the A operand
            {
                unsigned char *a_hi = smem + TcSmem::a(s, 0), *a_lo = smem + TcSmem::a(s, 1);   // this tile's half of the 128 rows
#pragma unroll
                for (int i = 0; i < kDGroups / 2; ++i)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int col = col0 + 8 * i + u;
                        if (col < frames) {
                            const float v = mag[2 * i + u], vh = tf32_trunc(v);
                            *reinterpret_cast<float *>(a_hi + a_off[u] + i * 1024) = vh;
                            *reinterpret_cast<float *>(a_lo + a_off[u] + i * 1024) = v - vh;
                            if constexpr (!kScaled) {
                                if (w.debug_band && in_band)
                                    w.debug_band[((int64_t)tw.ch * w.debug_cols + w.eval_offset + tw.e0 + tw.cols_before() + col) * L + bin] = v;
                            }
                        }
                    }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(a_ready);
            gcol += frames;
            if (gcol >= kStatRing) gcol -= kStatRing;
        }
        if ((it & 1) && lane == 0) ptx::mbar_arrive(a_ready);   // odd tile count: the last pair has no second tile
        tm.flush(dw == 0 && lane == 0);
    } else {
        // ================================ splitters (S) ================================================================
        const int st = (warp - kWarpS0) * 32 + lane;
        constexpr int kPerThread = kTileBytes / 16 / (kNumS * 32);
        static_assert(kPerThread * kNumS * 32 * 16 == kTileBytes, "tile size must divide over the splitters");
        TileWalk tw;
        tw.init(w, T);
        RoleTimer<kTiming> tm(w.debug_timing, 24);
        if constexpr (kDirect) {
            // Audio: global memory -> registers -> fp16 tile; no TMA, no raw fp32 tile in shared memory (which saves the TMA write
            // and the splitters' read of it: 544 of the ~2 600 shared-memory wavefronts a tile costs). A tile is 64 consecutive
            // hop-rows = one contiguous span of the channel; the six splitter warps take it as ONE batch: thread st loads the float4
            // st + 192 k (coalesced 512-byte warp loads, the whole tile in flight), converts to the two fp16 terms and stores 8 bytes
            // of each at the element's place in the K-major SWIZZLE_128B / _32B operand, then reloads the register with the same
            // element of the next tile. (All loads of a warp count on one scoreboard, so a two-batch pipeline inside a warp only
            // serialises; the overlap comes from the other roles' warps.) Warp 0 asks L2 for the tiles further ahead.
            // Instantiated per row length (hop / 4 = 32, 33 or 34 float4), which makes every slot count and predicate a constant.
            // kS16: the audio is 16-bit PCM (SYLDET_PCM_S16, planar): a slot is four samples = one 8-byte load, converted to k / 32768
            // exactly as ingest_kernel does - the conversion of SURVEY 8(f1) inside the first kernel, without the float32 round trip
            // through HBM (2 B in + 4 B out + 4 B re-read per sample become 2 B in).
            auto run = [&](auto r4c, auto s16c) {
                constexpr bool kS16 = decltype(s16c)::value;
                using Slot = std::conditional_t<kS16, uint2, float4>;
                constexpr int R4 = decltype(r4c)::value, kPerTile = kTileRows * R4, kSThreads = kNumSDirect * 32;
                constexpr int kFull = kPerTile / kSThreads, kRem = kPerTile % kSThreads, kSlots = kFull + (kRem ? 1 : 0);   // 10+1 | 11 | 11+1
                const bool has_last = kRem == 0 || st < kRem;   // this thread's last slot exists
                // Where this thread's float4 k lands in a tile buffer is the same for every tile. n < 128: fp16 chunk uu >> 4 (64 k
                // per 128-byte row), 16-byte unit (uu >> 1) & 7 swizzled by the row, half uu & 1, h2 two chunks further; n >= 128: the
                // SWIZZLE_32B tail, unit 0 = h1, unit 1 = h2, 16-byte unit ^= bit 2 of the row.
                uint32_t offp[kSlots];     // h1 offset | h2 offset << 16
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {
                    const int q = min(st + k * kSThreads, kPerTile - 1);
                    const int row = q / R4, uu = q - row * R4;
                    const uint32_t o1 = uu < 32 ? (uu >> 4) * kMainBytes + row * 128 + ((((uu >> 1) & 7) ^ (row & 7)) << 4) + (uu & 1) * 8
                                                : kMainChunks * kMainBytes + row * 32 + (((row >> 2) & 1) << 4) + (uu - 32) * 8;
                    const uint32_t o2 = uu < 32 ? o1 + 2 * kMainBytes : o1 ^ 16u;
                    offp[k] = o1 | (o2 << 16);
                }
                Slot v[kSlots];
                Slot zero_slot;
                if constexpr (kS16) zero_slot = make_uint2(0u, 0u);
                else zero_slot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < kSlots; ++k) v[k] = zero_slot;
                auto to_float4 = [](const Slot &r) {
                    if constexpr (kS16) {
                        const float sc = 1.0f / 32768.0f;
                        return make_float4((float)(short)(r.x & 0xFFFFu) * sc, (float)(short)(r.x >> 16) * sc, (float)(short)(r.y & 0xFFFFu) * sc,
                                           (float)(short)(r.y >> 16) * sc);
                    } else return r;
                };
                // Loads of the walk's tile. A tile whose 64 rows all exist (every one but the last of a channel) needs no predicates;
                // rows past the end of the channel read as zero.
                auto load_tile = [&]() {
                    const int64_t first = (int64_t)tw.ch * w.ch_stride + (int64_t)tw.first_row() * p.hop;   // in samples
                    const Slot *src;
                    if constexpr (kS16) src = reinterpret_cast<const Slot *>(w.pcm16 + first) + st;
                    else src = reinterpret_cast<const Slot *>(w.pcm + first) + st;
                    const int rows = w.n_rows - tw.first_row();
                    if (rows >= kTileRows) {
#pragma unroll
                        for (int k = 0; k < kSlots; ++k)
                            if (k < kFull || has_last) v[k] = __ldcs(src + k * kSThreads);
                    } else {
                        const int nq = max(rows, 0) * R4 - st;
#pragma unroll
                        for (int k = 0; k < kSlots; ++k) v[k] = k * kSThreads < nq ? __ldcs(src + k * kSThreads) : zero_slot;
                    }
                };
                int stg = 0;
                uint32_t stg_use = 0;
                if (tw.valid()) load_tile();
                while (tw.valid()) {
                    tm.wait(&lo_free[stg], (stg_use & 1) ^ 1, 1);   // the band DFT of the previous user of this buffer has read it
                    unsigned char *b16 = smem + stage_off(stg);
                    const long long t_c0 = tm.now();
                    uint2 hx[kSlots], hl[kSlots];
#pragma unroll
                    for (int k = 0; k < kSlots; ++k) {
                        split_fp16(to_float4(v[k]), hx[k], hl[k]);
                        if (k < kFull || has_last) {
                            *reinterpret_cast<uint2 *>(b16 + (offp[k] & 0xFFFFu)) = hx[k];
                            *reinterpret_cast<uint2 *>(b16 + (offp[k] >> 16)) = hl[k];
                        }
                    }
                    tw.next(w, T);
                    if (tw.valid()) load_tile();
                    tm.add(2, t_c0);
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&lo_ready[stg]);
                    if (++stg == n_stages) { stg = 0; ++stg_use; }
                }
            };
            if (w.pcm16) {
                if (p.hop == 132) run(std::integral_constant<int, 33>{}, std::true_type{});
                else if (p.hop == 128) run(std::integral_constant<int, 32>{}, std::true_type{});
                else run(std::integral_constant<int, 34>{}, std::true_type{});
            } else {
                if (p.hop == 132) run(std::integral_constant<int, 33>{}, std::false_type{});
                else if (p.hop == 128) run(std::integral_constant<int, 32>{}, std::false_type{});
                else run(std::integral_constant<int, 34>{}, std::false_type{});
            }
        } else
        for (uint32_t it = 0; tw.valid(); ++it, tw.next(w, T)) {
            const int s = it & 1;
            tm.wait(&full[s], (it >> 1) & 1, 0);
            const int ls = lo_stages == 2 ? s : 0;
            const uint32_t lo_use = lo_stages == 2 ? (it >> 1) : it;
            tm.wait(&lo_free[ls], (lo_use & 1) ^ 1, 1);  // pass 3 of the previous user of this lo buffer has read it
            if constexpr (kF16) {   // fp32 tile -> fp16 tile [h1 = fp16(x) | h2 = fp16((x - h1) * 2^11)]. A warp instruction takes rows r0, r0+4, r0+8, r0+12 (one 128 B row
                // per quarter warp): the loads are whole rows and the 8-byte stores of the four rows fall on disjoint banks.
                const unsigned char *hi_b = smem + TcSmem::hi(s);
                unsigned char *b16 = smem + (ls ? lo1_off : TcSmem::lo);
                const int sw = st >> 5, phys = lane & 7, rq = sw + 4 * (lane >> 3);   // row = rq + 16 * (kk & 3), chunk = kk >> 2
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
                        split_fp16(v[k], hx, hl);
                        *reinterpret_cast<uint2 *>(dst) = hx;
                        *reinterpret_cast<uint2 *>(dst + 2 * kMainBytes) = hl;
                    }
                }
                {   // tail: samples 128 .. 135 of row r (SWIZZLE_32B: 16-byte unit ^= bit 2 of the row)
                    const int r = st >> 1, ph = st & 1, flip = (r >> 2) & 1, u = ph ^ flip;
                    const float4 v = *reinterpret_cast<const float4 *>(hi_b + kMainChunks * kMainBytes + r * 32 + ph * 16);
                    unsigned char *dst = b16 + kMainChunks * kMainBytes + r * 32 + u * 8;
                    uint2 hx, hl;
                    split_fp16(v, hx, hl);
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
                    {
                        const float2 l01 = ptx::sub2(make_float2(v[k].x, v[k].y), make_float2(tf32_trunc(v[k].x), tf32_trunc(v[k].y)));
                        const float2 l23 = ptx::sub2(make_float2(v[k].z, v[k].w), make_float2(tf32_trunc(v[k].z), tf32_trunc(v[k].w)));
                        lo4[(b0 + k) * kNumS * 32] = make_float4(l01.x, l01.y, l23.x, l23.y);
                    }
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
// 16-bit PCM straight into the tensor kernel: the instantiation with the direct data path exists for the reference's sample shape
bool tc_direct_s16_supported(int hp, const FusedParams &p) {
    return hp == 4 && p.scaling == SYLDET_SCALING_LINEAR && p.window_stat == FUSED_STAT_L2 && p.tf[0] == SYLDET_TF_TANSIG && p.n_layers == 2 &&
           p.n_out == 1 && p.tf[1] == SYLDET_TF_PURELIN && p.n_op == 1;
}
int tc_k_pad() { return kKPad; }
int tc_a16_cols() { return kA16Cols; }
int tc_max_n0() { return kMaxN0; }
bool tc_layout_fits(int time_range, int n0) {
    // product ring: a pair (start b) reuses the positions of the previous pair except its last ring - pair rows, which this pair's
    // evaluations still read: ring - pair >= T - 1.
    // statistic ring: when the spectrum warps write tile it (even), the MMA warp has passed issue_l0(pair it/2 - 1), i.e. the
    // evaluators have copied pair it/2 - 3 (tiles it-6, it-5) and may still read the T-1 columns before it; the spectrum warps can
    // run two more tiles ahead before the next such wait: 9 tiles + T columns.
    return n0 <= kMaxN0 && kPairFrames + time_range - 1 <= kPRing && 9 * kTileFrames + time_range <= kStatRing;
}

cudaError_t launch_tc(int hp, int grid, size_t smem, const FusedParams &p, const TcWork &w, const void *tmap_main, const void *tmap_tail,
                      cudaStream_t stream) {
    const CUtensorMap &tm = *static_cast<const CUtensorMap *>(tmap_main);
    const CUtensorMap &tt = *static_cast<const CUtensorMap *>(tmap_tail);
    cudaError_t e = cudaErrorInvalidValue;
    int threads = kTcThreads;
    auto go = [&](auto kern) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kern<<<grid, threads, smem, stream>>>(p, w, tm, tt);
    };
    const bool scaled = p.scaling != SYLDET_SCALING_LINEAR;
    const bool fast = hp == 4 && !scaled && p.window_stat == FUSED_STAT_L2 && p.tf[0] == SYLDET_TF_TANSIG && p.n_layers == 2 && p.n_out == 1 &&
                      p.tf[1] == SYLDET_TF_PURELIN && p.n_op == 1;
    // the fp16 band DFT needs a per-window normaliser for its range guard (l2normalize or a min / max over the window) and linear
    // spectrogram scaling (a logarithm amplifies the errors of near-empty bins)
    const bool f16 = w.f16_corr && !scaled && (p.window_stat == FUSED_STAT_L2 || p.window_stat == FUSED_STAT_MINMAX);
    const bool direct = f16 && fast && (w.direct || w.pcm16 != nullptr);
    if (w.pcm16 && !direct) return cudaErrorNotSupported;   // the caller checks tc_direct_s16_supported first
    if (w.debug_timing) {   // SYLDET_TC_TIMING: instrumented build of the common shape only
        if (direct) { threads = kTcThreadsDirect; go(tc_detect_kernel<4, false, true, true, true, true>); }
        else if (f16 && fast) go(tc_detect_kernel<4, false, true, true, true, false>);
        else if (fast) go(tc_detect_kernel<4, false, true, true, false, false>);
        else return cudaErrorNotSupported;
    } else if (direct) { threads = kTcThreadsDirect; go(tc_detect_kernel<4, false, false, true, true, true>); }
    else if (f16 && fast) go(tc_detect_kernel<4, false, false, true, true, false>);
    else if (fast) go(tc_detect_kernel<4, false, false, true, false, false>);
    else if (f16 && hp == 4) go(tc_detect_kernel<4, false, false, false, true, false>);
    else if (f16) go(tc_detect_kernel<8, false, false, false, true, false>);
    else if (hp == 4 && !scaled) go(tc_detect_kernel<4, false, false, false, false, false>);
    else if (hp == 4) go(tc_detect_kernel<4, true, false, false, false, false>);
    else if (!scaled) go(tc_detect_kernel<8, false, false, false, false, false>);
    else go(tc_detect_kernel<8, true, false, false, false, false>);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace syldet
