// Launch wrappers for the CUDA kernels of the detection path.
#pragma once
#include <algorithm>

#include "device_types.cuh"
#include "../../include/syldet.h"

namespace syldet {

// ---- generic (reference-order) path: kernels_generic.cu ----------------------------------------------------------
cudaError_t launch_ingest(const void *src, int format, int interleaved, int n_channels, int64_t n_samples, int64_t src_stride,
                          float *dst, int64_t dst_stride, cudaStream_t stream);
cudaError_t launch_simulator_trace(const float *all_out, int n_channels, int64_t evals, int n_out, float thr0, int64_t first, int hop,
                                   int64_t n_samples, int format, void *trace, int64_t trace_stride, cudaStream_t stream);
// feat: [channel][n_cols][band] with `feat_ch_pitch` floats between channels; scaling_override >= 0 replaces the configuration's
// spectrogram scaling (SYLDET_SCALING_LINEAR = raw extractPower() values)
cudaError_t launch_stft_band_generic(const DevNet *d_net, int fft_len, const float *pcm, int64_t ch_stride, int n_channels,
                                     int64_t col0, int64_t n_cols, float *feat, int64_t feat_ch_pitch, int scaling_override,
                                     cudaStream_t stream);
cudaError_t launch_nn_generic(const DevNet *d_net, int max_width, const float *feat, int n_channels, int64_t n_cols,
                              int64_t n_evals, int64_t eval0, int64_t evals_total, int detect_rule, float *all_out,
                              EventSink sink, cudaStream_t stream);

// ---- live tick: kernels_generic.cu (stream_tick_kernel) ------------------------------------------------------------
enum { STREAM_PHASE_COPY = 1, STREAM_PHASE_COLUMNS = 2, STREAM_PHASE_EVALS = 4, STREAM_PHASE_ALL = 7 };
struct StreamTick {
    const float *staged;   // pinned host memory (device-visible) [n_channels][stage_pitch]: samples not yet on the device
    int stage_pitch, n_staged;
    float *ring;           // device [n_channels][ring_mask + 1] circular sample history
    int64_t ring_mask;     // capacity - 1 (power of two)
    int64_t ring_pos;      // absolute index of the first staged sample
    float *band;           // device [n_channels][band_mask + 1][band] circular band-feature columns
    int64_t band_mask;
    int64_t col0, n_cols;  // STFT columns this tick completes
    int64_t eval0, n_evals;  // evaluations this tick completes (evaluation j reads columns j .. j+T-1)
    float *out;            // pinned host memory (device-visible) [n_channels][n_evals][outputs]
    unsigned *counter;     // device, blocks finished (launches with several blocks per channel)
    unsigned *flags;       // pinned host memory [n_channels]: receive `seq` when the channel's results are visible (nullptr: no signal)
    uint4 *packed;         // pinned host memory [n_channels] or nullptr: {out0, out1, out2, seq} in one store (n_evals == 1, outputs <= 3)
    unsigned seq;
    int phases;            // STREAM_PHASE_* mask; COPY|COLUMNS|EVALS together need one block per channel
    const unsigned char *blob;  // the configuration's constant blob (DeviceModel): staged in shared memory when it fits
    int blob_bytes;        // multiple of 16; 0 = read constants through L2
    int work_bytes;        // set by launch_stream_tick: per-warp work area in front of the staged blob
    unsigned long long *level_in;  // device [n_channels]: max over buffers of the mean square (bits of a non-negative double); nullptr = off
    int *level_out;        // device [n_channels]: max over evaluations of output 0 (order-preserving int image of the float)
    int n_marks;           // staged buffers in this launch (<= kStreamMaxMarks)
    int marks[8];          // end of each staged buffer, in samples from the start of the staging area
    long long *stamps;     // optional pinned host [10]: clock64 at the phase boundaries of block (0,0) (SYLDET_STREAM_TIMING=1)
    // ResamplerLinear inside the tick (Processor.swift:116-121: resampleVector per channel before appendAudioData). The staged
    // samples are at the DEVICE rate; every staged buffer b is one resampleVector call producing rs_n_out[b] samples at ring
    // position ring_pos + rs_out0[b]. The data-independent phase (`offset`) comes from the host, `last` lives on the device.
    int rs_on;
    float rs_step;
    const float *rs_last_in;   // device [n_channels]: last sample of the previous tick's final buffer (Resampler.swift:66)
    float *rs_last_out;        // device [n_channels]: written by this tick (the two arrays alternate, so a launch never reads what it writes)
    float rs_offset[8];
    int rs_n_out[8];
    int rs_out0[8];
};
constexpr size_t kStreamTickMaxSmem = 200 * 1024;
constexpr int kStreamMaxMarks = 8;
size_t stream_tick_smem(int fft_len, int max_width, int *warps_out);
cudaError_t launch_stream_tick(const DevNet *d_net, int fft_len, int max_width, int n_channels, int blocks_per_channel,
                               StreamTick t, cudaStream_t stream);

// ---- fused fast path: kernels_fused.cu ---------------------------------------------------------------------------
constexpr int kFusedMaxW0 = 6144;     // folded layer-0 weights, floats ([inputs][HP])
constexpr int kFusedMaxHidden = 8;    // widest layer the register epilogue handles
constexpr int kFusedMaxLayers = 4;
constexpr int kFusedMaxOut = 8;
constexpr int kFusedThreads = 256;
constexpr int kFusedMaxBand = 128;    // band bins per column (4 per lane)

enum { FUSED_STAT_NONE = 0, FUSED_STAT_L2 = 1, FUSED_STAT_MINMAX = 2, FUSED_STAT_STD = 3 };

// Passed by value in kernel parameter space (constant bank): the weights are warp-uniform operands of the FFMAs.
struct alignas(16) FusedParams {
    int win_len, gap, hop, k0, band, time_range, scaling;
    int band_pitch;   // floats between ring columns (odd => conflict-free column reads)
    int ring_cols;    // ring capacity in columns
    int nn_tile;      // evaluations that trigger an epilogue pass
    int abuf_floats;  // one audio staging buffer, floats (multiple of 4)
    int window_stat;  // FUSED_STAT_*
    int n_layers, n_out, n_op, reserved0;
    int tf[kFusedMaxLayers];
    int width[kFusedMaxLayers];     // outputs of each layer
    float v[kFusedMaxHidden];       // V_h  = sum_i W_hi A_i
    float bprime[kFusedMaxHidden];  // B'_h = b_h + sum_i W_hi C_i
    float rest_w[(kFusedMaxLayers - 1) * kFusedMaxHidden * kFusedMaxHidden];  // layers >= 1, [out][in] padded to 8x8
    float rest_b[(kFusedMaxLayers - 1) * kFusedMaxHidden];
    float op_y[kMaxProcessing];
    float op_gain[kMaxProcessing * kFusedMaxOut];
    float op_xoff[kMaxProcessing * kFusedMaxOut];
    double thr[kFusedMaxOut];
    float thr_f[kFusedMaxOut];      // smallest float >= thr: (double)v >= thr  <=>  v >= thr_f for every float v
    alignas(16) float w0[kFusedMaxW0];  // [inputs][HP]: W_hi * A_i, hidden index fastest
};

struct FusedWork {
    const float *pcm;      // planar float32
    const float *pcm_begin, *pcm_end;  // valid range of the whole buffer (bounds for 16-byte bulk copies)
    int64_t ch_stride;
    int n_channels;
    int64_t evals_per_channel;
    int64_t chunk_evals;
    int chunks_per_channel;
    int detect_rule;       // SYLDET_DETECT_*
    int64_t eval_offset;   // index of the first evaluation this launch handles (events, outputs and audio are offset by it)
    int64_t out_evals_per_channel;  // evaluations per channel in all_out (its channel pitch)
    float *all_out;        // optional [n_channels][out_evals_per_channel][n_out]
    EventSink sink;
    const float *window;   // [win_len]
    const float2 *twiddle; // [fft_len/2]
    float *debug_band;     // optional [n_channels][debug_cols][band] band magnitudes before the scaling (tests / spectra API)
    int64_t debug_cols;
};

struct FusedLaunch {
    int fft_len = 0, hp = 0;
    int grid = 0;
    size_t smem = 0;
};

// Geometry helpers shared by host planning and the kernel.
__host__ __device__ constexpr int fused_r1(int fft_len) { return fft_len >= 256 ? 16 : 8; }
__host__ __device__ constexpr int fused_r2(int fft_len) { return (fft_len / 2) / fused_r1(fft_len); }
__host__ __device__ constexpr int fused_group(int fft_len) { return 32 * fused_r1(fft_len) / (fft_len / 2); }  // frames per warp pass
__host__ __device__ constexpr int fused_round_cols(int fft_len) { return (kFusedThreads / 32) * fused_group(fft_len); }
__host__ __device__ constexpr int fused_frame_pitch(int fft_len) { return fft_len / 2 + (fft_len / 2) / fused_r1(fft_len); }  // float2 per frame

// ---- tensor-core variant: kernels_tc.cu ---------------------------------------------------------------------------
struct TcWork {
    int n_channels;
    int chunks_per_channel;
    int64_t evals_per_channel;      // evaluations this launch handles per channel (all of their rows are complete)
    int64_t chunk_evals;
    int64_t eval_offset, out_evals_per_channel;
    int detect_rule;
    float *all_out;
    EventSink sink;
    const float *dft_hi, *dft_lo;   // [128][k_pad] windowed DFT matrix, tf32 hi / lo parts
    const float *wcat_hi, *wcat_lo; // [n0][32] folded layer-0 weights, row (t*HP + h), column = band bin
    const uint32_t *dft16;          // [128][tc_a16_cols()] fp16 pairs: the two-term split of the DFT matrix (kF16 variant), see plan_tc
    int n0;                         // T*HP rounded up to a multiple of 16
    int lo_stages;                  // 1 or 2 lo tiles in shared memory (tc_lo_stages)
    int f16_corr;                   // 1: band DFT on fp16 splits (sample shape; range-guarded); 0: 3xTF32 (SYLDET_KERNEL_TENSOR_TF32, SYLDET_TC_TF32_CORR=1)
    float *debug_band;              // optional [n_channels][debug_cols][band] band magnitudes (tests)
    int64_t debug_cols;
    long long *debug_timing;        // optional [grid][32] cycle counters per role (SYLDET_TC_TIMING=1)
    // range guard of the fp16 correction pass (f16_corr = 1): an evaluation whose window energy sum |X|^2 is outside
    // [guard_lo, FLT_MAX] (and not exactly 0) sets *range_flag; the host then repeats the launch with f16_corr = 0
    float guard_lo;
    float guard_range;              // same guard for windows normalised by their min / max: the smallest window range max - min
    int *range_flag;
    // direct variant of the fp16 band DFT (f16_corr = 1): the splitter warps read the audio from global memory themselves
    // (no TMA, no raw fp32 tile in shared memory)
    int direct;
    int pf_dist;                    // tiles ahead of the splitters that are requested into L2 (0 = no prefetch)
    const float *pcm;               // first sample of evaluation eval_offset of channel 0
    const int16_t *pcm16;           // the same position in a planar 16-bit PCM buffer: the direct path converts k / 32768 itself (else nullptr)
    int64_t ch_stride;              // floats between channels
    int n_rows;                     // complete hop-rows per channel from `pcm` on
    int zero;                       // 0 (an operand the compiler cannot fold: orders the splitters' loads behind their arrival waits)
};
size_t tc_smem_bytes(const FusedParams &p, int hp);
int tc_lo_stages(const FusedParams &p, int hp);
int tc_tile_frames();
int tc_plan_unit_tiles(int64_t eval_count, int n_channels, int resident, int tile_frames, int warm);   // engine.cu: tiles per unit of a launch
bool tc_direct_s16_supported(int hp, const FusedParams &p);
int tc_k_pad();
int tc_a16_cols();                  // 32-bit words per row of the fp16 A operand (TcWork::dft16)
int tc_max_n0();
bool tc_layout_fits(int time_range, int n0);
cudaError_t launch_tc(int hp, int grid, size_t smem, const FusedParams &p, const TcWork &w, const void *tmap_main,
                      const void *tmap_tail, cudaStream_t stream);

// ---- wide-hidden tensor path: kernels_wide.cu (+ stft_planes_kernel in kernels_generic.cu) -----------------------------------
constexpr int kWidePL = 6;           // planes (of 4 bins) per K chunk: 24 inputs per (chunk, column offset) step
constexpr int kWideStftCols = 64;    // STFT columns per CTA of stft_planes_kernel (they share one staged audio span)

// Passed by value in kernel parameter space.
struct alignas(16) WideParams {
    int time_range, band, n_planes;  // n_planes = ceil(band / 4) rounded up to a multiple of kWidePL
    int hidden, h_pad;               // h_pad = hidden rounded up to a multiple of 256 (accumulator passes of 256 columns)
    int n_out, n_op, window_stat;    // FUSED_STAT_NONE / _L2 / _MINMAX
    int tf0, tf1, reserved0, reserved1;
    float b1[kFusedMaxOut];          // output-layer biases
    float op_y[kMaxProcessing];
    float op_gain[kMaxProcessing * kFusedMaxOut];
    float op_xoff[kMaxProcessing * kFusedMaxOut];
    float thr_f[kFusedMaxOut];       // smallest float >= threshold
};

struct WideWork {
    const float *planes_hi, *planes_lo;  // [n_channels][n_planes][rows_alloc][4]: band magnitudes, raw and (v - tf32(v))
    const float4 *stats;                 // [n_channels][rows_alloc]: per column {sum x^2, min, max, -}
    int64_t rows_alloc;                  // row pitch of the planes / statistics
    int64_t n_cols;                      // magnitude rows that exist (n_evals + T - 1)
    int64_t n_evals;                     // evaluations of this launch per channel
    int n_channels;
    int detect_rule;
    int64_t eval_offset, out_evals_per_channel;
    float *all_out;
    EventSink sink;
    const float *weights;                // pre-arranged blocks, see plan_wide: [pass of 256 hidden units][chunk][t][hi | lo][plane][256][4]
    const float *v, *bprime;             // [h_pad] folded layer-0 constants: z = acc * alpha + beta * V + B'
    const float *w1;                     // [n_out][h_pad] output-layer weights
};
size_t wide_smem_bytes(int h_pad);
int wide_tile_rows();
int wide_max_time_range();
size_t wide_weight_block_bytes();
cudaError_t launch_wide(int grid, const WideParams &p, const WideWork &w, cudaStream_t stream);
bool stft_planes_fast_supported(int fft_len);
cudaError_t launch_stft_planes_fast(const DevNet *d_net, int fft_len, int win_len, int band, int hop, const float *pcm, int64_t ch_stride,
                                    int n_channels, int64_t col0, int64_t n_cols, float *hi, float *lo, float4 *stats, int n_planes,
                                    int64_t rows_alloc, cudaStream_t stream);
size_t stft_planes_smem(int fft_len, int win_len, int hop, int n_planes);
cudaError_t launch_stft_planes(const DevNet *d_net, int fft_len, int win_len, int hop, const float *pcm, int64_t ch_stride, int n_channels,
                               int64_t col0, int64_t n_cols, float *hi, float *lo, float4 *stats, int n_planes, int64_t rows_alloc,
                               cudaStream_t stream);

bool fused_supports_fft(int fft_len);
size_t fused_smem_bytes(int fft_len, const FusedParams &p);
cudaError_t launch_fused(const FusedLaunch &cfg, const FusedParams &p, const FusedWork &w, cudaStream_t stream);
cudaError_t fused_max_blocks_per_sm(int fft_len, int hp, size_t smem, int *blocks);
// live tick on the register FFT and the folded network (kernels_fused.cu: stream_tick_fast_kernel); one block per channel, all phases;
// d_params: device copy of the FusedParams
// geometry of one tick in shared memory (host: tick_geom in kernels_fused.cu)
struct TickGeom {
    int win_len, hop, gap, k0, band, time_range;
    int span_floats;   // samples from the start of the first new frame to the end of the last one, rounded up to 4
    int win_floats;    // band values from the oldest column of the first evaluation window to the newest column, rounded up to 4
    int stage_floats;  // staged samples, rounded up to 4
};

bool stream_tick_fast_supported(int fft_len, const FusedParams &p);
bool stream_tick_fast_fits(int fft_len, int hp, const FusedParams &p, const StreamTick &t);   // this tick's shared-memory footprint is within the cap
cudaError_t launch_stream_tick_fast(int fft_len, int hp, const FusedParams &host_params, const FusedParams *d_params, const StreamTick &t,
                                    const float *window, const float2 *twiddle, int n_channels, cudaStream_t stream);
// resident tick (stream_tick_resident_kernel): blocks stay on their SMs; a dispatcher block polls the message the host posts in pinned
// memory and republishes it in a device mailbox of stream_tick_post_bytes() bytes (+ one `leave` word)
size_t stream_tick_post_bytes();
void stream_tick_post_write(void *post, const StreamTick &t);   // what changes per tick, as quads {3 words, t.seq}, the first one last
bool stream_tick_resident_plan(int fft_len, int hp, const FusedParams &p, int stage_cap, TickGeom *gmax);
bool stream_tick_resident_tick_fits(const TickGeom &gmax, const FusedParams &p, const StreamTick &t);
cudaError_t launch_stream_tick_resident(int fft_len, int hp, const FusedParams *d_params, const void *post, const unsigned *quit, unsigned *alive,
                                        void *mailbox, unsigned *leave, const StreamTick &first, const TickGeom &gmax, long long idle_cycles,
                                        const float *window, const float2 *twiddle, int n_channels, int sm_count, cudaStream_t stream);

}  // namespace syldet
