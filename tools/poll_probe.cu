// Round trip host -> resident kernel -> host through pinned memory: the host posts a message of Q 16-byte quads {payload, seq}, B blocks
// of T threads poll it (every thread its quad, or thread 0 only), then thread 0 of each block answers with one store (+ optional
// system fence). Prints the median / p99 round trip in microseconds for a few shapes. Build: nvcc -O2 -arch=sm_100a -o poll_probe poll_probe.cu
#include <cuda_runtime.h>
#include <emmintrin.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint4 ldv4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ldv(const unsigned *p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// mode 0: every thread polls its quad, block-wide vote (the shape of stream_tick_resident_kernel); mode 1: thread 0 polls quad 0, then
// the block reads the rest; mode 2: thread 0 polls, nobody else reads
__global__ void probe(const uint4 *post, int quads, unsigned *answer, int rounds, int mode, int fence, long long max_cycles) {
    const int tid = threadIdx.x;
    const long long t0 = clock64();
    for (unsigned k = 1; k <= (unsigned)rounds; ++k) {
        if (mode == 0) {
            const int q = tid < quads ? tid : 0;
            for (;;) {
                const uint4 v = ldv4(post + q);
                const bool stop = tid == 0 && clock64() - t0 > max_cycles;
                if (__syncthreads_or(stop)) return;
                if (__syncthreads_and(v.w == k)) break;
            }
        } else {
            __shared__ int go;
            if (tid == 0) {
                go = 1;
                while (ldv4(post).w != k)
                    if (clock64() - t0 > max_cycles) { go = 0; break; }
            }
            __syncthreads();
            if (!go) return;
            if (mode == 1 && tid < quads) {
                const uint4 v = ldv4(post + tid);
                if (v.w != k) asm volatile("trap;");
            }
            __syncthreads();
        }
        if (tid == 0) {
            *(volatile unsigned *)(answer + blockIdx.x) = k;
            if (fence) __threadfence_system();
        }
    }
}

int main() {
    uint4 *post;
    unsigned *answer;
    const int max_quads = 64, max_blocks = 148;
    cudaMallocHost(&post, max_quads * sizeof(uint4));
    cudaMallocHost(&answer, max_blocks * sizeof(unsigned));
    const int rounds = 2000;
    struct Shape { int blocks, threads, quads, mode, fence; };
    const Shape shapes[] = {{1, 32, 1, 2, 0},  {1, 32, 1, 2, 1},   {64, 32, 1, 2, 1},  {64, 128, 37, 1, 1}, {64, 128, 37, 0, 1},
                            {64, 128, 37, 0, 0}, {8, 128, 37, 0, 1}, {1, 128, 37, 0, 1}, {64, 64, 37, 0, 1},  {148, 128, 37, 0, 1}};
    for (const Shape &s : shapes) {
        std::memset(post, 0, max_quads * sizeof(uint4));
        std::memset(answer, 0, max_blocks * sizeof(unsigned));
        probe<<<s.blocks, s.threads>>>(post, s.quads, answer, rounds, s.mode, s.fence, 4000000000LL);   // ~2 s of cycles at most
        std::vector<double> us;
        bool dead = false;
        for (unsigned k = 1; k <= (unsigned)rounds && !dead; ++k) {
            const auto a = std::chrono::steady_clock::now();
            for (int q = s.quads - 1; q >= 0; --q) _mm_store_si128((__m128i *)(post + q), _mm_set_epi32((int)k, q, q, q));
            _mm_sfence();
            for (int b = 0; b < s.blocks; ++b) {
                long spins = 0;
                while (((volatile unsigned *)answer)[b] != k)
                    if (++spins > 400000000L) { dead = true; break; }
                if (dead) break;
            }
            us.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count());
        }
        cudaError_t e = cudaDeviceSynchronize();
        std::sort(us.begin(), us.end());
        std::printf("blocks %3d threads %3d quads %2d mode %d fence %d: p50 %.2f us, p99 %.2f us%s (%s)\n", s.blocks, s.threads, s.quads, s.mode, s.fence,
                    us.empty() ? 0.0 : us[us.size() / 2], us.empty() ? 0.0 : us[us.size() * 99 / 100], dead ? " TIMED OUT" : "", cudaGetErrorString(e));
    }
    return 0;
}
