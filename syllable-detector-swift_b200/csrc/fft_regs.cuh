// In-register DFTs of size 2, 4, 8, 16 (natural order in and out) and their constexpr roots of unity: the butterflies of the
// shared-memory Stockham transforms in kernels_fused.cu (N <= 512, two passes) and kernels_wide.cu (stft_planes_fast_kernel).
#pragma once
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"

namespace syldet {
namespace {

// ---- constexpr roots of unity: w_R^k = exp(-2 pi i k / R), R | 16 --------------------------------------------------
__host__ __device__ constexpr double cos16(int i) {  // cos(2 pi i / 16)
    i = ((i % 16) + 16) % 16;
    switch (i) {
        case 0: return 1.0;
        case 1: case 15: return 0.92387953251128673848;
        case 2: case 14: return 0.70710678118654752440;
        case 3: case 13: return 0.38268343236508977173;
        case 4: case 12: return 0.0;
        case 5: case 11: return -0.38268343236508977173;
        case 6: case 10: return -0.70710678118654752440;
        case 7: case 9: return -0.92387953251128673848;
        default: return -1.0;
    }
}
__host__ __device__ constexpr double sin16(int i) { return cos16(i - 4); }

// Complex add / subtract: a (re, im) pair is one aligned register pair, so each is ONE packed instruction (FADD2, sm_100) with
// the rounding of the two scalar adds. SYLDET_FFT_SCALAR_ADDS keeps the scalar form (kernels built with -fmad=false semantics do
// not care: an add is an add).
#ifdef SYLDET_FFT_SCALAR_ADDS
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return ptx::add2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return ptx::sub2(a, b); }
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// a * w_R^K with the trivial cases folded
template <int R, int K>
__device__ __forceinline__ float2 mul_root(float2 a) {
    constexpr int i = K * (16 / R);  // sixteenths of a turn
    if constexpr (i == 0) return a;
    else if constexpr (i == 4) return make_float2(a.y, -a.x);  // * (-i)
    else if constexpr (i == 2) {
        constexpr float c = (float)0.70710678118654752440;
        return make_float2((a.x + a.y) * c, (a.y - a.x) * c);
    } else if constexpr (i == 6) {
        constexpr float c = (float)0.70710678118654752440;
        return make_float2((a.y - a.x) * c, -(a.x + a.y) * c);
    } else {
        constexpr float cr = (float)cos16(i), ci = (float)(-sin16(i));
        return make_float2(a.x * cr - a.y * ci, a.x * ci + a.y * cr);
    }
}

// In-register DFT, natural order in and out: v[k] <- sum_n v[n] exp(-2 pi i n k / R)
template <int R>
struct Dft {
    template <int K>
    static __device__ __forceinline__ void combine(float2 *v, const float2 *e, const float2 *o) {
        if constexpr (K < R / 2) {
            const float2 t = mul_root<R, K>(o[K]);
            v[K] = cadd(e[K], t);
            v[K + R / 2] = csub(e[K], t);
            combine<K + 1>(v, e, o);
        }
    }
    static __device__ __forceinline__ void run(float2 *v) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        Dft<R / 2>::run(e);
        Dft<R / 2>::run(o);
        combine<0>(v, e, o);
    }
};
template <>
struct Dft<2> {
    static __device__ __forceinline__ void run(float2 *v) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <>
struct Dft<1> {
    static __device__ __forceinline__ void run(float2 *) {}
};

}  // namespace
}  // namespace syldet
