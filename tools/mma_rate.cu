// Microbenchmark: cycles per tcgen05.mma for the shapes the tensor-core kernel uses (run on the GPU box).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I syllable-detector-swift_b200/csrc tools/mma_rate.cu -o gpurun_out/mma_rate
// One CTA per SM (all SMs busy, like the real kernel); one thread issues `reps` MMAs, commits, waits; clock64 around it.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "ptx_sm100.cuh"

using namespace syldet;

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// kind::f16, bf16 x bf16 -> f32
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}

// K-major, no swizzle (canonical ((8,n),2):((1,SBO),LBO)): the layout of the wide kernel's sliding A operand and weight planes
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

// mode: 0 tf32 TS, 1 tf32 SS, 2 bf16 TS, 3 bf16 SS, 4 tf32 SS with no-swizzle plane operands (A planes of 272 rows, B planes of 256 rows)
template <int mode, int m, int n, int chain>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long *out) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.0f;
    if (tid == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(&tmem_ptr, 512);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    {   // zero the A region of TMEM (columns 0..127)
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 128; c += 8) ptx::tmem_st_x8(tmem + ((uint32_t)(warp * 32) << 16) + c, z);
        ptx::tc_wait_st();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 0 && elect_one()) {
        const uint32_t a_s = ptx::smem_addr(smem), b_s = ptx::smem_addr(smem + 32768);
        const uint32_t idesc = (mode < 2 || mode == 4) ? ptx::idesc_tf32(m, n) : idesc_bf16(m, n);
        const uint32_t d0 = tmem + 256;
        uint32_t phase = 0;
        // warm-up
        for (int r = 0; r < 8; ++r) {
            if (mode == 0) ptx::mma_tf32_ts(d0, tmem, ptx::smem_desc_kmajor(b_s, 1024, 2), idesc, 0);
            if (mode == 1) ptx::mma_tf32_ss(d0, ptx::smem_desc_kmajor(a_s, 1024, 2), ptx::smem_desc_kmajor(b_s, 1024, 2), idesc, 0);
            if (mode == 2) mma_f16_ts(d0, tmem, ptx::smem_desc_kmajor(b_s, 1024, 2), idesc, 0);
            if (mode == 3) mma_f16_ss(d0, ptx::smem_desc_kmajor(a_s, 1024, 2), ptx::smem_desc_kmajor(b_s, 1024, 2), idesc, 0);
        }
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, phase);
        phase ^= 1;
        const long long t0 = clock64();
#pragma unroll 8
        for (int r = 0; r < reps; ++r) {
            // chain = 1: every MMA accumulates into the same D (like a K loop); chain = 0: alternate two accumulators
            const uint32_t d = d0 + ((chain || !(r & 1)) ? 0 : 128);
            const uint32_t ko = (r & 3) * 32;  // walk the 4 K slices of a 128-byte swizzle row
            if (mode == 0) ptx::mma_tf32_ts(d, tmem + (r & 3) * 8, ptx::smem_desc_kmajor(b_s + ko, 1024, 2), idesc, 1);
            if (mode == 1) ptx::mma_tf32_ss(d, ptx::smem_desc_kmajor(a_s + ko, 1024, 2), ptx::smem_desc_kmajor(b_s + ko, 1024, 2), idesc, 1);
            if (mode == 2) mma_f16_ts(d, tmem + (r & 3) * 8, ptx::smem_desc_kmajor(b_s + ko, 1024, 2), idesc, 1);
            if (mode == 3) mma_f16_ss(d, ptx::smem_desc_kmajor(a_s + ko, 1024, 2), ptx::smem_desc_kmajor(b_s + ko, 1024, 2), idesc, 1);
            if (mode == 4) ptx::mma_tf32_ss(d, desc_nosw(a_s + (r & 7) * 16, 4352, 128), desc_nosw(b_s, 4096, 128), idesc, 1);
        }
        const long long t1 = clock64();
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, phase);
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

template <int mode, int m, int n, int chain>
void run(long long *d_out) {
    const char *names[5] = {"tf32 TS", "tf32 SS", "bf16 TS", "bf16 SS", "tf32 SS no-swizzle planes"};
    const int reps = 2048;
    cudaFuncSetAttribute(rate_kernel<mode, m, n, chain>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    rate_kernel<mode, m, n, chain><<<148, 128, 100 * 1024>>>(reps, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s m=%d n=%d: %s\n", names[mode], m, n, cudaGetErrorString(e));
        exit(1);
    }
    long long h[2];
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    printf("%s M=%3d N=%3d %s: issue %.1f cyc/mma, complete %.1f cyc/mma\n", names[mode], m, n, chain ? "same-D " : "alt-D  ",
           (double)h[0] / reps, (double)h[1] / reps);
}

template <int mode>
void run_mode(long long *d_out) {
    run<mode, 128, 16, 1>(d_out);
    run<mode, 128, 48, 1>(d_out);
    run<mode, 128, 64, 1>(d_out);
    run<mode, 128, 64, 0>(d_out);
    run<mode, 128, 128, 1>(d_out);
    run<mode, 128, 256, 1>(d_out);
    run<mode, 64, 8, 1>(d_out);
    run<mode, 64, 16, 1>(d_out);
    run<mode, 64, 48, 1>(d_out);
    run<mode, 64, 48, 0>(d_out);
    run<mode, 64, 64, 1>(d_out);
    run<mode, 64, 128, 1>(d_out);
    run<mode, 64, 256, 1>(d_out);
}

// Whole-chip throughput of one shape over a long launch, by wall clock (CUDA events): the denominator of the wide-hidden kernel's
// roofline (kind::tf32, M = 128, N = 256, both operands from shared memory), next to the bf16 figure MEASURED_PEAKS.json holds.
template <int mode>
double peak_tflops(long long *d_out, int reps, float *ms_out) {
    cudaFuncSetAttribute(rate_kernel<mode, 128, 256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    rate_kernel<mode, 128, 256, 1><<<148, 128, 100 * 1024>>>(reps / 8, d_out);   // warm-up
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    rate_kernel<mode, 128, 256, 1><<<148, 128, 100 * 1024>>>(reps, d_out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms;
    const double k = mode < 2 ? 8.0 : 16.0;
    return 148.0 * reps * 2.0 * 128.0 * 256.0 * k / (ms * 1e-3) / 1e12;
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--nosw") {
        long long *d_out;
        cudaMalloc(&d_out, 16);
        run<1, 128, 256, 1>(d_out);
        run<4, 128, 256, 1>(d_out);
        run<4, 128, 128, 1>(d_out);
        run<4, 128, 64, 1>(d_out);
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "--peak") {   // JSON line: burst (~20 ms) and sustained (~2 s) dense peaks
        long long *d_out;
        cudaMalloc(&d_out, 16);
        float ms;
        const double tf32_burst = peak_tflops<1>(d_out, 200000, &ms);
        const double tf32_sust = peak_tflops<1>(d_out, 20000000, &ms);
        const float tf32_ms = ms;
        const double bf16_burst = peak_tflops<3>(d_out, 200000, &ms);
        const double bf16_sust = peak_tflops<3>(d_out, 20000000, &ms);
        printf("{\"tf32_tflops\": %.1f, \"tf32_tflops_sustained\": %.1f, \"bf16_tflops\": %.1f, \"bf16_tflops_sustained\": %.1f, "
               "\"sustained_ms\": %.0f, \"how\": \"tools/mma_rate.cu --peak: tcgen05.mma cta_group::1 M=128 N=256 (K=8 tf32 / K=16 bf16), "
               "operands from shared memory, one accumulator chain per CTA, 148 CTAs, CUDA events\"}\n",
               tf32_burst, tf32_sust, bf16_burst, bf16_sust, tf32_ms);
        return 0;
    }

    long long *d_out;
    cudaMalloc(&d_out, 16);
    run_mode<0>(d_out);
    run_mode<1>(d_out);
    run_mode<2>(d_out);
    run_mode<3>(d_out);
    return 0;
}
