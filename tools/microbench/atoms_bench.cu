// Micro-benchmark: shared-memory atomic throughput of the table layouts considered for kmer_hist_kernel.
//   nvcc -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o /tmp/atoms_bench tools/microbench/atoms_bench.cu
//   /tmp/atoms_bench [iterations]
// Every warp owns a private table and posts pseudo-random updates; prints lane-updates per clock per SM and clocks per
// warp-wide instruction.  Round 2 added the cost-model cases: exact conflict degrees, partial warps, one address, 64-bit
// adds and the non-atomic load/add/store alternative.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void red_inc(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }
__device__ __forceinline__ void red_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add64(uint32_t a, unsigned long long v) { asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t xs(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

enum {
    M_POPC = 0,        // POPC.INC, random words of a WORDS-word table
    M_PACK16 = 1,      // packed 16-bit counters (variable increment), random words
    M_LANE_ADD = 2,    // lane-private words (bank = lane), variable increment: conflict-free ATOMS.ADD
    M_THIRD = 3,       // like 0, lanes with lane % 3 == 0 only
    M_LANE_INC = 4,    // lane-private words, constant 1: conflict-free ATOMS.POPC.INC
    M_Q8 = 5,          // like 0, lanes 0..7 only
    M_H16 = 6,         // like 0, lanes 0..15 only
    M_SAME = 7,        // every lane the same word
    M_LDST = 8,        // non-atomic: lane-private ld.shared + add + st.shared
    M_WAY2 = 9,        // exactly two lanes per bank, different words
    M_WAY4 = 10,       // exactly four lanes per bank, different words
    M_ADD64 = 11,      // lane-private 64-bit adds (two banks per lane: 2-way by construction)
    M_Q8_FREE = 12,    // lanes 0..7 only, conflict-free
};

template <int MODE, int WORDS, int PER_ITER>
__global__ void __launch_bounds__(256) bench(int iters, uint32_t *sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tab = smem_u32(sm) + warp * WORDS * 4;
    for (int i = lane; i < WORDS; i += 32) reinterpret_cast<uint32_t *>(sm)[warp * WORDS + i] = 0;
    __syncwarp();
    uint32_t s = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
    constexpr uint32_t ROWS = WORDS / 32;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < PER_ITER; j += 2) {
            const uint32_t r = xs(s);
            const uint32_t r0 = r & (WORDS - 1), r1 = (r >> 16) & (WORDS - 1);
            if (MODE == M_POPC) {
                red_inc(tab + (r0 << 2));
                red_inc(tab + (r1 << 2));
            } else if (MODE == M_PACK16) {
                red_add(tab + (r0 << 2), (r & 0x8000u) ? 65536u : 1u);
                red_add(tab + (r1 << 2), (r & 0x80000000u) ? 65536u : 1u);
            } else if (MODE == M_LANE_ADD) {
                red_add(tab + ((r & (ROWS - 1)) << 7) + (lane << 2), 1u << ((r >> 6) & 24));
                red_add(tab + (((r >> 16) & (ROWS - 1)) << 7) + (lane << 2), 1u << ((r >> 22) & 24));
            } else if (MODE == M_THIRD) {
                if (lane % 3 == 0) red_inc(tab + (r0 << 2));
                if (lane % 3 == 0) red_inc(tab + (r1 << 2));
            } else if (MODE == M_LANE_INC) {
                red_inc(tab + ((r & (ROWS - 1)) << 7) + (lane << 2));
                red_inc(tab + (((r >> 16) & (ROWS - 1)) << 7) + (lane << 2));
            } else if (MODE == M_Q8) {
                if (lane < 8) red_inc(tab + (r0 << 2));
                if (lane < 8) red_inc(tab + (r1 << 2));
            } else if (MODE == M_H16) {
                if (lane < 16) red_inc(tab + (r0 << 2));
                if (lane < 16) red_inc(tab + (r1 << 2));
            } else if (MODE == M_SAME) {
                const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, r0, 0), w1 = __shfl_sync(0xFFFFFFFFu, r1, 0);
                red_inc(tab + (w0 << 2));
                red_inc(tab + (w1 << 2));
            } else if (MODE == M_LDST) {
                const uint32_t a0 = tab + ((r & (ROWS - 1)) << 7) + (lane << 2), a1 = tab + (((r >> 16) & (ROWS - 1)) << 7) + (lane << 2);
                sts(a0, lds(a0) + 1u);
                sts(a1, lds(a1) + 1u);
            } else if (MODE == M_WAY2) {       // bank = lane >> 1 (+16 for the second use), rows differ inside a pair
                red_inc(tab + ((((r & (ROWS / 2 - 1)) << 1) | (lane & 1)) << 7) + ((lane >> 1) << 2));
                red_inc(tab + (((((r >> 16) & (ROWS / 2 - 1)) << 1) | (lane & 1)) << 7) + (((lane >> 1) + 16) << 2));
            } else if (MODE == M_WAY4) {
                red_inc(tab + ((((r & (ROWS / 4 - 1)) << 2) | (lane & 3)) << 7) + ((lane >> 2) << 2));
                red_inc(tab + (((((r >> 16) & (ROWS / 4 - 1)) << 2) | (lane & 3)) << 7) + (((lane >> 2) + 8) << 2));
            } else if (MODE == M_ADD64) {
                red_add64(tab + ((r & (ROWS / 2 - 1)) << 8) + (lane << 3), (r & 0x8000u) ? (1ull << 32) : 1ull);
                red_add64(tab + (((r >> 16) & (ROWS / 2 - 1)) << 8) + (lane << 3), (r & 0x80000000u) ? (1ull << 32) : 1ull);
            } else if (MODE == M_Q8_FREE) {
                if (lane < 8) red_inc(tab + ((r & (ROWS - 1)) << 7) + (lane << 2));
                if (lane < 8) red_inc(tab + (((r >> 16) & (ROWS - 1)) << 7) + (lane << 2));
            }
        }
    }
    __syncwarp();
    uint32_t acc = 0;
    for (int i = lane; i < WORDS; i += 32) acc += reinterpret_cast<uint32_t *>(sm)[warp * WORDS + i];
    if (acc == 0xdeadbeef) sink[0] = acc;
}

static int g_iters = 20000;

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
    const int iters = g_iters;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<sms * ctas_per_sm, 256, smem>>>(100, sink);
    cudaEventRecord(a);
    k<<<sms * ctas_per_sm, 256, smem>>>(iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double updates = (double)ctas_per_sm * 8 * 32 * lanes_frac * (double)iters * PER_ITER;   // per SM
    const double clk = ms * 1e-3 * khz * 1e3;
    printf("%-52s warps/SM %3d  %8.3f ms  lane-updates/clk/SM %6.2f  clk per warp-instruction %5.2f  err=%s\n", name, ctas_per_sm * 8, ms,
           updates / clk, clk / ((double)ctas_per_sm * 8 * iters * PER_ITER), cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main(int argc, char **argv) {
    if (argc > 1) g_iters = atoi(argv[1]);
    run<M_POPC, 1024, 8>("popc.inc 1024 words (5-mer table)", 5);
    run<M_POPC, 1024, 8>("popc.inc 1024 words, 2 CTAs", 2);
    run<M_POPC, 4096, 8>("popc.inc 4096 words (6-mer, 32-bit)", 1);
    run<M_POPC, 256, 8>("popc.inc 256 words (4-mer table)", 5);
    run<M_PACK16, 2048, 8>("packed u16 add 2048 words (6-mer, 16-bit)", 3);
    run<M_PACK16, 512, 8>("packed u16 add 512 words (5-mer, 16-bit)", 5);
    run<M_LANE_ADD, 2048, 8>("lane-private words, add (conflict-free)", 3);
    run<M_LANE_INC, 2048, 8>("lane-private words, popc.inc (conflict-free)", 3);
    run<M_WAY2, 2048, 8>("two lanes per bank (2-way conflict)", 3);
    run<M_WAY4, 2048, 8>("four lanes per bank (4-way conflict)", 3);
    run<M_SAME, 1024, 8>("one word for the whole warp", 5);
    run<M_THIRD, 1024, 8>("popc.inc 1024 words, a third of the lanes", 5, 11.0 / 32.0);
    run<M_H16, 1024, 8>("popc.inc 1024 words, lanes 0..15", 5, 0.5);
    run<M_Q8, 1024, 8>("popc.inc 1024 words, lanes 0..7", 5, 0.25);
    run<M_Q8_FREE, 2048, 8>("lane-private words, lanes 0..7 (conflict-free)", 3, 0.25);
    run<M_ADD64, 2048, 8>("lane-private 64-bit add", 3);
    run<M_LDST, 2048, 8>("non-atomic lane-private ld + add + st", 3);
    return 0;
}
