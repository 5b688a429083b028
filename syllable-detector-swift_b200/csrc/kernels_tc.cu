// Tensor-core variant of the fused detection kernel (sm_100a): the band DFT runs on tcgen05 as a 3xTF32 contraction.
//
// Same path and citations as kernels_fused.cu (CSTFT.swift:280-337, SyllableDetector.swift:134-217, NeuralNet.swift:294-326,
// TrackDetector.swift:71-77); only the transform differs:
//   * the audio of a channel is viewed as a row-major matrix Y[row][hop] (row r = samples [r*hop, (r+1)*hop)); a frame that
//     starts at row c covers row c and the first W-hop samples of row c+1, so
//         X_c[k] = sum_n Y[c][n] B1[n][k] + sum_n Y[c+1][n] B2[n][k],   B1/B2 = the two halves of the windowed DFT matrix;
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B) lands 64-row tiles of Y in shared memory as the K-major B operand;
//   * the DFT matrix [128 x K] (rows: Re B1 | Im B1 | Re B2 | Im B2, 32 bins each) lives in TMEM as the A operand, split
//     into tf32 hi + lo; the audio lo part (x - tf32(x)) is produced by the worker warps; three MMA passes
//     (Ahi*Bhi + Ahi*Blo + Alo*Bhi) accumulate D[128 x 64] in TMEM (FP32), double buffered;
//   * worker warps read D with tcgen05.ld, add the row-shifted halves, take magnitudes of the band bins into the
//     shared-memory ring, and run the same per-evaluation epilogue as the SIMT kernel.
// Roles: warps 0-7 workers (TMEM lane quadrant = warp % 4), warp 8 TMA producer, warp 9 MMA issuer + TMEM allocator.
#include <cuda.h>

#include "fused_epilogue.cuh"
#include "ptx_sm100.cuh"

namespace syldet {

namespace {

constexpr int kTcWorkers = 8;                  // worker warps
constexpr int kTcThreads = (kTcWorkers + 2) * 32;
constexpr int kTileRows = 64;                  // rows of Y per tile = N of the MMA
constexpr int kTileFrames = kTileRows - 1;     // frames completed per tile (the last row only feeds the previous frame)
constexpr int kMainChunks = 4;                 // 32-float K chunks (SWIZZLE_128B)
constexpr int kTailCols = 8;                   // remaining K columns (SWIZZLE_32B)
constexpr int kKPad = kMainChunks * 32 + kTailCols;  // 136
constexpr int kMainBytes = kTileRows * 128;    // one main chunk
constexpr int kTailBytes = kTileRows * 32;
constexpr int kTileBytes = kMainChunks * kMainBytes + kTailBytes;  // 34 816 per hi (or lo) tile
constexpr int kTmemCols = 512;
constexpr int kColAhi = 0, kColAlo = kKPad, kColD0 = 2 * kKPad;   // D buffers: 64 columns each
static_assert(kColD0 + 2 * kTileRows <= kTmemCols, "TMEM budget");

struct TcSmem {  // offsets from the 1024-byte aligned base
    __host__ __device__ static constexpr int tile(int stage, int lo) { return (stage * 2 + lo) * 35840; }  // 34 816 rounded up to 1024
    static constexpr int xbuf = 4 * 35840;                        // [4 quadrants][64 frames][32 bins] float
    static constexpr int bars = xbuf + 4 * kTileRows * 32 * 4;    // 10 mbarriers + tmem pointer
    static constexpr int utw_unused = bars + 128;
    static constexpr int ring = utw_unused;                       // band-magnitude ring
};

// Linear walk over (unit, tile) pairs owned by this CTA; every role iterates the identical sequence.
struct TileWalk {
    int64_t unit, n_units;
    int tile, ntiles, ch, ncols;
    int64_t e0;
    __device__ void load(const TcWork &w, int T) {
        if (unit >= n_units) return;
        ch = (int)(unit / w.chunks_per_channel);
        e0 = (unit - (int64_t)ch * w.chunks_per_channel) * w.chunk_evals;
        const int ne = (int)min(w.chunk_evals, w.evals_per_channel - e0);
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
};

template <int HP>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_detect_kernel(const __grid_constant__ FusedParams p, const TcWork w, const __grid_constant__ CUtensorMap tmap_main,
                 const __grid_constant__ CUtensorMap tmap_tail) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TcSmem::bars);
    uint64_t *full = bars, *ready = bars + 2, *stage_free = bars + 4, *tmem_full = bars + 6, *tmem_empty = bars + 8;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 10);
    float *xbuf = reinterpret_cast<float *>(smem + TcSmem::xbuf);
    float *ring = reinterpret_cast<float *>(smem + TcSmem::ring);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.band, T = p.time_range;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&full[i], 1);
            ptx::mbar_init(&ready[i], kTcWorkers);
            ptx::mbar_init(&stage_free[i], 1);
            ptx::mbar_init(&tmem_full[i], 1);
            ptx::mbar_init(&tmem_empty[i], kTcWorkers);
        }
        ptx::fence_mbar_init();
    }
    if (warp == kTcWorkers + 1) {
        ptx::tmem_alloc(tmem_ptr, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // DFT matrix -> TMEM (A operand): lane m = matrix row, column = k. Rows are split hi | lo.
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
                const uint32_t ph = (it >> 1) & 1;
                ptx::mbar_wait(&stage_free[s], ph ^ 1);  // first use of each stage passes immediately
                unsigned char *dst = smem + TcSmem::tile(s, 0);
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
            constexpr uint32_t idesc = ptx::idesc_tf32(128, kTileRows);
            TileWalk tw;
            tw.init(w, T);
            for (uint32_t it = 0; tw.valid(); ++it, tw.next(w, T)) {
                const int s = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                ptx::mbar_wait(&ready[s], ph);            // hi landed (TMA) and lo written (workers)
                ptx::mbar_wait(&tmem_empty[s], ph ^ 1);   // accumulator s drained by the epilogue
                ptx::tc_fence_after();
                const uint32_t d = tmem_base + kColD0 + s * kTileRows;
                const uint32_t hi = ptx::smem_addr(smem + TcSmem::tile(s, 0)), lo = ptx::smem_addr(smem + TcSmem::tile(s, 1));
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a = tmem_base + (pass == 2 ? kColAlo : kColAhi);
                    const uint32_t b = pass == 1 ? lo : hi;
#pragma unroll
                    for (int j = 0; j < kMainChunks; ++j)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            ptx::mma_tf32_ts(d, a + j * 32 + ks * 8, ptx::smem_desc_kmajor(b + j * kMainBytes + ks * 32, 1024, 2), idesc, acc);
                            acc = 1;
                        }
                    ptx::mma_tf32_ts(d, a + kMainChunks * 32, ptx::smem_desc_kmajor(b + kMainChunks * kMainBytes, 256, 6), idesc, acc);
                }
                ptx::mma_commit(&tmem_full[s]);
                ptx::mma_commit(&stage_free[s]);
            }
        }
    } else {
        // ================================ workers =====================================================================
        const int quad = warp & 3, half = warp >> 2;
        const int wtid = tid;  // 0..255
        auto worker_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(kTcWorkers * 32) : "memory"); };
        auto split_lo = [&](uint32_t it) {  // lo = x - tf32_trunc(x) for the whole tile (layout-agnostic: same offsets in both buffers)
            const int s = it & 1;
            ptx::mbar_wait(&full[s], (it >> 1) & 1);
            const float4 *hi4 = reinterpret_cast<const float4 *>(smem + TcSmem::tile(s, 0));
            float4 *lo4 = reinterpret_cast<float4 *>(smem + TcSmem::tile(s, 1));
            for (int i = wtid; i < kTileBytes / 16; i += kTcWorkers * 32) {
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

        TileWalk cur;
        cur.init(w, T);
        if (cur.valid()) split_lo(0);
        int cols_done = 0, evals_done = 0, col_slot = 0, eval_slot = 0;
        for (uint32_t it = 0; cur.valid(); ++it) {
            TileWalk nxt = cur;
            nxt.next(w, T);
            if (nxt.valid()) split_lo(it + 1);  // overlaps the MMAs of tile `it`

            const int s = it & 1;
            if (cur.tile == 0) { cols_done = evals_done = col_slot = eval_slot = 0; }
            const int frames = min(kTileFrames, cur.ncols - cur.tile * kTileFrames);  // frames this tile completes

            // ---- D (TMEM) -> xbuf[quadrant][frame][bin] ----------------------------------------------------------
            ptx::mbar_wait(&tmem_full[s], (it >> 1) & 1);
            ptx::tc_fence_after();
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
            // ---- X_c = (Re1[c] + Re2[c+1]) + i (Im1[c] + Im2[c+1]); |X| for the band -> ring -------------------------
            for (int c = warp; c < frames; c += kTcWorkers) {
                int slot = col_slot + c;
                if (slot >= p.ring_cols) slot -= p.ring_cols;
                const float re = xbuf[(0 * kTileRows + c) * 32 + lane] + xbuf[(2 * kTileRows + c + 1) * 32 + lane];
                const float im = xbuf[(1 * kTileRows + c) * 32 + lane] + xbuf[(3 * kTileRows + c + 1) * 32 + lane];
                float mag = sqrt_fast(re * re + im * im);
                if (p.scaling != SYLDET_SCALING_LINEAR) mag = scale_value(mag, p.scaling);
                if (lane < L) ring[slot * p.band_pitch + lane] = mag;
                if (w.debug_band && lane < L)
                    w.debug_band[((int64_t)cur.ch * w.debug_cols + cur.e0 + cur.tile * kTileFrames + c) * L + lane] = mag;
            }
            worker_sync();
            cols_done += frames;
            col_slot += frames;
            if (col_slot >= p.ring_cols) col_slot -= p.ring_cols;

            // ---- per-evaluation epilogue ---------------------------------------------------------------------------------
            const bool last_tile = cur.tile == cur.ntiles - 1;
            const int n_ready = cols_done - (T - 1) - evals_done;
            if (n_ready >= p.nn_tile || (last_tile && n_ready > 0)) {
                float *out_base = w.all_out ? w.all_out + ((int64_t)cur.ch * w.out_evals_per_channel + w.eval_offset + cur.e0) * p.n_out : nullptr;
                for (int qb = warp * 32; qb < n_ready; qb += kTcWorkers * 32) {
                    const int q = qb + lane;
                    float out[kFusedMaxOut];
                    bool hit = false;
                    if (q < n_ready) {
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
                                w.sink.events[idx] = DevEvent{cur.ch, 0, w.eval_offset + cur.e0 + evals_done + q};
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
                if (last_tile) worker_sync();  // the next unit restarts the ring
            }
            cur = nxt;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kTcWorkers + 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

size_t tc_smem_bytes(const FusedParams &p) {
    return 1024 + TcSmem::ring + (size_t)p.ring_cols * p.band_pitch * sizeof(float);
}

int tc_tile_frames() { return kTileFrames; }
int tc_k_pad() { return kKPad; }

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
