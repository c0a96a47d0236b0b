// Micro-benchmark: shared-memory atomic throughput of the table layouts considered for kmer_hist_kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/atoms_bench tools/microbench/atoms_bench.cu
// Every warp owns a private table and posts pseudo-random updates; prints lane-updates per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void red_inc(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
__device__ __forceinline__ void red_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t xs(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// MODE 0: POPC.INC, WORDS-word table.  MODE 1: packed 16-bit counters (variable increment) in a WORDS-word table.
// MODE 2: lane-private byte counters, 64 rows x 32 lanes.  MODE 3: like 0 but only lanes with (lane % 3 == 0) post (a third of a warp).
template <int MODE, int WORDS, int PER_ITER>
__global__ void __launch_bounds__(256) bench(int iters, uint32_t *sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tab = smem_u32(sm) + warp * WORDS * 4;
    for (int i = lane; i < WORDS; i += 32) reinterpret_cast<uint32_t *>(sm)[warp * WORDS + i] = 0;
    __syncwarp();
    uint32_t s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < PER_ITER; j += 2) {
            const uint32_t r = xs(s);
            if (MODE == 0) {
                red_inc(tab + ((r & (WORDS - 1)) << 2));
                red_inc(tab + (((r >> 16) & (WORDS - 1)) << 2));
            } else if (MODE == 1) {
                red_add(tab + ((r & (WORDS - 1)) << 2), (r & 0x8000u) ? 65536u : 1u);
                red_add(tab + (((r >> 16) & (WORDS - 1)) << 2), (r & 0x80000000u) ? 65536u : 1u);
            } else if (MODE == 2) {
                red_add(tab + ((r & 63) << 7) + (lane << 2), 1u << ((r >> 6) & 24));
                red_add(tab + (((r >> 16) & 63) << 7) + (lane << 2), 1u << ((r >> 22) & 24));
            } else {
                if (lane % 3 == 0) red_inc(tab + ((r & (WORDS - 1)) << 2));
                if (lane % 3 == 0) red_inc(tab + (((r >> 16) & (WORDS - 1)) << 2));
            }
        }
    }
    __syncwarp();
    uint32_t acc = 0;
    for (int i = lane; i < WORDS; i += 32) acc += reinterpret_cast<uint32_t *>(sm)[warp * WORDS + i];
    if (acc == 0xdeadbeef) sink[0] = acc;
}

template <int MODE, int WORDS, int PER_ITER>
static void run(const char *name, int ctas_per_sm, double lanes_frac = 1.0) {
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const size_t smem = 8 * WORDS * 4;
    auto k = bench<MODE, WORDS, PER_ITER>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 256, smem);
    if (ctas_per_sm > per_sm) ctas_per_sm = per_sm;
    uint32_t *sink; cudaMalloc(&sink, 4);
    const int iters = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<sms * ctas_per_sm, 256, smem>>>(100, sink);
    cudaEventRecord(a);
    k<<<sms * ctas_per_sm, 256, smem>>>(iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double updates = (double)ctas_per_sm * 8 * 32 * lanes_frac * (double)iters * PER_ITER;   // per SM
    const double clk = ms * 1e-3 * khz * 1e3;
    printf("%-44s warps/SM %3d  %.3f ms  lane-updates/clk/SM %.2f  clk per warp-atomic %.2f  err=%s\n", name, ctas_per_sm * 8, ms,
           updates / clk, clk / ((double)ctas_per_sm * 8 * iters * PER_ITER), cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    run<0, 1024, 8>("popc.inc 1024 words (today, 5-mer table)", 5);
    run<0, 1024, 8>("popc.inc 1024 words, 2 CTAs", 2);
    run<0, 4096, 8>("popc.inc 4096 words (6-mer, 32-bit)", 1);
    run<0, 256, 8>("popc.inc 256 words (4-mer table)", 5);
    run<1, 2048, 8>("packed u16 add 2048 words (6-mer, 16-bit)", 3);
    run<1, 2048, 8>("packed u16 add 2048 words, 2 CTAs", 2);
    run<1, 512, 8>("packed u16 add 512 words (5-mer, 16-bit)", 5);
    run<2, 2048, 8>("lane-private byte counters (conflict-free)", 3);
    run<3, 1024, 8>("popc.inc 1024 words, a third of the lanes", 5, 11.0 / 32.0);
    return 0;
}
