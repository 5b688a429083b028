// FP32 FFMA peak of the box (SURVEY.md section 8d asks for it beside the HBM peak): every thread runs 16 independent FMA chains,
// 148 SMs x 8 CTAs x 256 threads, wall clock by CUDA events. Prints one JSON line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ffma_peak.cu -o gpurun_out/ffma_peak && gpurun_out/ffma_peak
#include <cstdio>

__global__ void __launch_bounds__(256) ffma_kernel(float *out, int iters, float a, float b) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;   // never true: keeps the chains alive
}

static double run(int iters, float *d, float *ms_out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    ffma_kernel<<<148 * 8, 256>>>(d, iters / 8, 0.999f, 1e-3f);
    cudaEventRecord(a);
    ffma_kernel<<<148 * 8, 256>>>(d, iters, 0.999f, 1e-3f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms;
    return 148.0 * 8 * 256 * (double)iters * 8 * 16 * 2 / (ms * 1e-3) / 1e12;
}

int main() {
    float *d;
    cudaMalloc(&d, 4);
    float ms1, ms2;
    const double burst = run(20000, d, &ms1), sustained = run(2000000, d, &ms2);
    printf("{\"fp32_ffma_tflops\": %.2f, \"fp32_ffma_tflops_sustained\": %.2f, \"burst_ms\": %.1f, \"sustained_ms\": %.0f, "
           "\"how\": \"tools/ffma_peak.cu: 16 independent FFMA chains per thread, 148 x 8 CTAs x 256 threads, CUDA events\"}\n",
           burst, sustained, ms1, ms2);
    return cudaGetLastError() != cudaSuccess;
}
